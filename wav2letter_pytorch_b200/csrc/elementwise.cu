// Memory-bound companions of the conv kernels: layout changes, BatchNorm statistics / apply / backward with
// fused activation, dropout, reflection halo and length mask, log_softmax fwd/bwd, bias gradient, casts.
// All activations are time-major [B, T, C] bf16; every kernel moves 16 bytes (8 channels) per thread access.
#include <string.h>

#include "common.cuh"

namespace w2l {

// device-resident dropout epoch registered by the host (w2l_set_dropout_epoch); per process, like the GEMM scratch
static const uint64_t* g_dropout_epoch = nullptr;

// ---------------------------------------------------------------- dropout keep-bits
// 32 keep-bits at a time -- the 4 consecutive rows x 8 channels a thread of the BatchNorm passes stages together: bit 8u + i belongs to
// row 4*rg + u, channel 8*cv + i -- each 1 with probability keep_q / 2^kDropBits.  Bitwise Bernoulli synthesis: walking the binary digits
// of the keep probability from the least significant one, m <- digit ? (m | w) : (m & w) with a fresh uniform word w per digit gives
// P(bit) <- (digit + P(bit)) / 2, i.e. the keep probability resolved to 2^-12 for all 32 elements from 12 hash words; the words are
// murmur3's 32-bit finaliser over a counter keyed by (seed, row group, channel vector).  ~35 instructions per 16-byte vector where
// the Philox4x32-10 stream of the first cut spent ~115 (one call per 8 elements, 16 bits per element).  The realised keep
// probability keep_q / 4096 -- not the requested 1 - p, which differs from it by < 1.3e-4 -- scales the kept elements, so the
// expectation is exact (nn.Dropout semantics, wav2letter.py:44 / jasper.py:372-376).  The forward pass stores the bits (drop_mask)
// and the backward passes read them back.
constexpr int kDropBits = 12;
__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}
__device__ __forceinline__ uint32_t dropout_mask32(uint64_t seed, uint32_t rg, uint32_t cv, uint32_t keep_q) {
  const uint32_t base = fmix32(fmix32((uint32_t)seed ^ rg) + (uint32_t)(seed >> 32) + cv * 0x9E3779B1u);
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < kDropBits; ++j) {
    const uint32_t w = fmix32(base + (uint32_t)(j + 1) * 0x9E3779B9u);
    const uint32_t digit = 0u - ((keep_q >> j) & 1u);                // all ones / all zeros
    m = (m & w) | (digit & (m ^ w));                               // digit ? (m | w) : (m & w): one LOP3
  }
  return m;
}
// The seed a launch draws its keep-bits from: the by-value seed of the call, advanced by the device-resident epoch when one is
// registered -- a captured CUDA graph replays the same by-value arguments, the epoch (bumped inside the graph) makes every replay
// draw a fresh mask while forward and backward of one step still agree.
template <typename Args>
__device__ __forceinline__ uint64_t drop_seed(const Args& a) {
  return a.epoch ? a.seed + *a.epoch * 0xA0761D6478BD642Full : a.seed;
}

__device__ __forceinline__ void unpack8(const uint4& q, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 q;
  q.x = pack_bf16x2(v[0], v[1]);
  q.y = pack_bf16x2(v[2], v[3]);
  q.z = pack_bf16x2(v[4], v[5]);
  q.w = pack_bf16x2(v[6], v[7]);
  return q;
}
// scalar store of one value into bf16 / fp32 activation storage
__device__ __forceinline__ void store_act(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ void store_act(float* p, float v) { *p = v; }
__device__ __forceinline__ float load_act(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ float load_act(const float* p) { return *p; }

// 8 consecutive channels of one row as a thread stages them: one 16-byte vector of bf16, or two of fp32 (the fp32-faithful mode,
// W2L_STORE_F32: activations stay fp32 between the tf32 GEMMs).  Every BatchNorm / activation pass is written once over this.
template <typename TA>
struct Vec8;
template <>
struct Vec8<__nv_bfloat16> {
  uint4 q;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { q = __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ __forceinline__ void unpack(float (&v)[8]) const { unpack8(q, v); }
  __device__ __forceinline__ void pack(const float (&v)[8]) { q = pack8(v); }
  __device__ __forceinline__ void zero() { q = make_uint4(0u, 0u, 0u, 0u); }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const { *reinterpret_cast<uint4*>(p) = q; }
};
template <>
struct Vec8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = __ldg(reinterpret_cast<const float4*>(p));
    b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  }
  __device__ __forceinline__ void unpack(float (&v)[8]) const {
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  __device__ __forceinline__ void pack(const float (&v)[8]) {
    a = make_float4(v[0], v[1], v[2], v[3]);
    b = make_float4(v[4], v[5], v[6], v[7]);
  }
  __device__ __forceinline__ void zero() { a = b = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = a;
    *(reinterpret_cast<float4*>(p) + 1) = b;
  }
};
__device__ __forceinline__ void load8f(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// ---------------------------------------------------------------- NCW fp32 -> (unfolded, padded) time-major bf16
// block: 32 output rows of one utterance; the needed input span is staged (transposed) in shared memory.
template <typename TOut>
__global__ void im2col_ncw_kernel(const float* __restrict__ x, TOut* __restrict__ out, int F, int T, int rows, int k,
                                  int stride, int dil, int pad_left, int pad_mode, const int32_t* __restrict__ lens, int span,
                                  int pitch) {
  extern __shared__ float tile[];   // [F][pitch]
  const int b = blockIdx.y, r0 = blockIdx.x * 32;
  const int nrows = min(32, rows - r0);
  const int tstart = r0 * stride - pad_left;
  const int len = lens ? min(T, max(0, lens[b])) : T;
  const float* xb = x + (int64_t)b * F * T;
  for (int i = threadIdx.x; i < F * span; i += blockDim.x) {
    const int f = i / span, dt = i - f * span;
    int t = tstart + dt;
    float v = 0.f;
    if (pad_mode == W2L_PAD_REFLECT) {
      if (t < 0) t = -t;
      if (t >= T) t = 2 * (T - 1) - t;
      if (t >= 0 && t < T) v = xb[(int64_t)f * T + t];
    } else if (t >= 0 && t < T) {
      v = xb[(int64_t)f * T + t];
    }
    if (t >= len) v = 0.f;
    tile[f * pitch + dt] = v;
  }
  __syncthreads();
  const int KF = k * F;
  TOut* ob = out + ((int64_t)b * rows + r0) * KF;
  if ((F & 7) == 0 && sizeof(TOut) == 2) {   // 16-byte stores: one thread packs 8 consecutive features of one (row, tap)
    const int F8 = F >> 3, KF8 = k * F8;
    for (int i = threadIdx.x; i < nrows * KF8; i += blockDim.x) {
      const int r = i / KF8, c = i - r * KF8;
      const int j = c / F8, f0 = (c - j * F8) << 3;
      const float* src = tile + f0 * pitch + r * stride + j * dil;
      uint4 q;
      q.x = pack_bf16x2(src[0], src[pitch]);
      q.y = pack_bf16x2(src[2 * pitch], src[3 * pitch]);
      q.z = pack_bf16x2(src[4 * pitch], src[5 * pitch]);
      q.w = pack_bf16x2(src[6 * pitch], src[7 * pitch]);
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ob) + (int64_t)r * KF + j * F + f0) = q;
    }
    return;
  }
  for (int i = threadIdx.x; i < nrows * KF; i += blockDim.x) {
    const int r = i / KF, c = i - r * KF;
    const int j = c / F, f = c - j * F;
    store_act(ob + i, tile[f * pitch + r * stride + j * dil]);
  }
}

// ---------------------------------------------------------------- time-major unfold / fold (strided layers beyond the first)
// out[b, t, j*C + c] = x[b, t*stride + j*dil - pad_left, c] (zero outside [0, x_rows)): the strided conv becomes a k=1 GEMM over
// out.  One thread moves 16 bytes (8 channels) of one (row, tap).
__global__ void im2col_tm_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int B, int x_rows, int C,
                                 int T_out, int k, int stride, int dil, int pad_left) {
  const int c8 = C >> 3;
  const int64_t total = (int64_t)B * T_out * k * c8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c8) * 8;
    int64_t q = i / c8;
    const int j = (int)(q % k);
    q /= k;
    const int t = (int)(q % T_out), b = (int)(q / T_out);
    const int r = t * stride + j * dil - pad_left;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r >= 0 && r < x_rows) v = __ldg(reinterpret_cast<const uint4*>(x + ((int64_t)b * x_rows + r) * C + c));
    *reinterpret_cast<uint4*>(out + (((int64_t)b * T_out + t) * k + j) * C + c) = v;
  }
}
// The adjoint (gather form, fp32 accumulation, no atomics): dx[b, r, c] = sum over taps j with (r + pad_left - j*dil) = t*stride,
// 0 <= t < T_out, of dcol[b, t, j*C + c].
__global__ void col2im_tm_kernel(const __nv_bfloat16* __restrict__ dcol, __nv_bfloat16* __restrict__ dx, int B, int x_rows, int C,
                                 int T_out, int k, int stride, int dil, int pad_left) {
  const int c8 = C >> 3;
  const int64_t total = (int64_t)B * x_rows * c8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c8) * 8;
    const int64_t q = i / c8;
    const int r = (int)(q % x_rows), b = (int)(q / x_rows);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < k; ++j) {
      const int u = r + pad_left - j * dil;
      if (u < 0) break;                                   // u only decreases with j
      const int t = u / stride;
      if (t * stride != u || t >= T_out) continue;
      float v[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(dcol + (((int64_t)b * T_out + t) * k + j) * C + c)), v);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += v[e];
    }
    *reinterpret_cast<uint4*>(dx + ((int64_t)b * x_rows + r) * C + c) = pack8(acc);
  }
}

// time-major (bf16 | f32) -> NCW fp32, 32x32 shared-memory transpose
template <typename TIn>
__global__ void tm_to_ncw_kernel(const TIn* __restrict__ x, float* __restrict__ out, int T, int C, int64_t x_batch_stride,
                                 int x_row_offset, int ld, int out_pitch) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const TIn* xb = x + (int64_t)b * x_batch_stride + (int64_t)x_row_offset * ld;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int t = t0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (t < T && c < C) ? (float)xb[(int64_t)t * ld + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, t = t0 + threadIdx.x;
    if (c < C && t < T) out[((int64_t)b * C + c) * out_pitch + t] = tile[threadIdx.x][i];
  }
}

// ---------------------------------------------------------------- BatchNorm statistics
// block (32, 8): 32 threads x 8 channels = 256 channels, 8 row lanes; grid (ceil(C/256), row_blocks)
__global__ void bn_stats_kernel(const __nv_bfloat16* __restrict__ z, int64_t rows, int C, float* __restrict__ stats,
                                int rows_per_block) {
  __shared__ float s_sum[8][256 + 8], s_sq[8][256 + 8];
  const int c = blockIdx.x * 256 + threadIdx.x * 8;
  const int64_t r_begin = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r_end = min(rows, r_begin + rows_per_block);
  float sum[8], sq[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) sum[i] = sq[i] = 0.f;
  if (c < C) {
    for (int64_t r = r_begin + threadIdx.y; r < r_end; r += 8) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(z + r * C + c));
      float v[8];
      unpack8(q, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        sum[i] += v[i];
        sq[i] = fmaf(v[i], v[i], sq[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s_sum[threadIdx.y][threadIdx.x * 8 + i] = sum[i];
    s_sq[threadIdx.y][threadIdx.x * 8 + i] = sq[i];
  }
  __syncthreads();
  const int tid = threadIdx.y * 32 + threadIdx.x;   // 256 threads -> one channel each
  const int cc = blockIdx.x * 256 + tid;
  if (cc < C) {
    float a = 0.f, q = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) {
      a += s_sum[y][tid];
      q += s_sq[y][tid];
    }
    atomicAdd(stats + cc, a);
    atomicAdd(stats + C + cc, q);
  }
}

__global__ void bn_finalize_kernel(const float* __restrict__ stats, int64_t rows, int C, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, const float* __restrict__ conv_bias, float eps, float momentum,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ mean_out, float* __restrict__ invstd_out,
                                   int64_t* __restrict__ num_batches_tracked) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && num_batches_tracked) *num_batches_tracked += 1;     // BatchNorm1d's counter buffer, without a launch of its own
  if (c >= C) return;
  const double n = (double)rows;
  const double mean = (double)stats[c] / n;
  double var = (double)stats[C + c] / n - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[c] : 1.f, bt = beta ? beta[c] : 0.f;
  scale[c] = g * invstd;
  shift[c] = bt - (float)mean * g * invstd;
  if (mean_out) mean_out[c] = (float)mean;
  if (invstd_out) invstd_out[c] = invstd;
  if (running_mean) {
    const float m_full = (float)mean + (conv_bias ? conv_bias[c] : 0.f);
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * m_full;
    const double unbiased = rows > 1 ? var * n / (n - 1.0) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// ---------------------------------------------------------------- BN apply + dropout + act + halo + mask
// Thread mapping of the three kernels below (round 2): a CTA is bx x ny threads with bx = min(C/8, 256) channel vectors (16 bytes =
// 8 bf16 channels each) and ny = 256 / bx row lanes.  For every width up to 2048 channels bx == C/8, so the threads of a CTA walk
// CONSECUTIVE 16-byte vectors of the row-major activation -- a warp's access is one contiguous run whatever C is (the first cut tiled
// the channels in blocks of 256: a 896-wide layer ran its fourth column block with half-empty warps) -- and a thread keeps its 8
// channels, hence their constants in registers, for the whole kernel.  blockIdx.y owns a contiguous range of (b, t) rows, a thread
// walks it in GROUPS OF 4 CONSECUTIVE ROWS (the unit the dropout bits are drawn for); the grid holds 2-3 CTAs per SM, each looping
// over hundreds of rows, so that the per-CTA fixed costs (constants, block reduction, one atomic per channel) are amortised.
// These passes are INSTRUCTION-issue bound before they are HBM bound -- at 6.5 TB/s a thread may spend ~170 issue slots per 32 bytes
// it moves, and the first cut spent ~250 (profiles/r2_bn_passes.md) -- hence: dropout bits for 32 elements from 14 32-bit hashes,
// dropout scaling folded into the per-channel constants, NaN-propagating min/max instead of compare+select, one integer division per
// row group, float rsqrt + one Newton step in the finalize fold.
constexpr int kBnThreads = 256;
constexpr int kBnFwdCtasPerSm = 3;   // forward pass: <= 80 registers
constexpr int kBnBwdCtasPerSm = 2;   // backward passes: 3 streams + 5 per-channel constants per thread want ~110 registers
constexpr int kRowGroup = 4;         // consecutive rows a thread stages together: all their 16-byte loads are issued before any is used

struct BnFwdArgs {
  const void* z;               // activations: bf16, or fp32 under W2L_STORE_F32 (the kernels' TA)
  const void* res;
  const float* scale;          // given affine (eval / separately finalised statistics); unused when stats != null
  const float* shift;
  const float* res_scale;
  const float* res_shift;
  // stats != null: the batch statistics the conv epilogue accumulated -> scale / shift here, in every CTA (a few flops per channel),
  // instead of a bn_finalize launch in between; CTA row 0 also writes what backward needs and moves the running statistics
  const float* stats;          // [2C] sum, sum of squares over stat_rows rows
  int64_t stat_rows;
  const float* gamma;
  const float* beta;
  const float* conv_bias;
  float eps, momentum;
  float* running_mean;
  float* running_var;
  int64_t* num_batches_tracked;
  float* fin;                  // [4][C] scale, shift, mean, invstd
  void* y;
  int B, T, C, pl, pr;
  uint32_t keep_q;             // dropout: keep probability in units of 2^-kDropBits (0: no dropout)
  float inv_keep;              // 2^kDropBits / keep_q
  uint64_t seed;
  const uint64_t* epoch;       // device-resident step counter folded into the seed (w2l_set_dropout_epoch), or nullptr
  const int32_t* lens;
  uint8_t* drop_mask;          // [B*T*C/8] keep-bits: written by the forward pass, read back by the backward passes
  float* zero_ptr;             // a small buffer this launch clears for a LATER kernel (the layer's backward reduction sums)
  int zero_count;
  int rows_per_block;
};

struct BnBwdArgs {
  const void* z;
  const void* res;
  const void* dyp;
  const float* scale;
  const float* shift;
  const float* res_scale;
  const float* res_shift;
  const float* mean;
  const float* invstd;
  const float* gamma;
  float* red;                  // [2C] sum g, sum g*xhat: accumulated by the reduce pass (zero on entry), read by the apply pass
  float* red_out;              // apply pass: copy of red for the caller (dbeta, dgamma), so that `red` itself can be recycled
  void* dz;
  void* g_out;
  int dz_rows;
  int B, T, C, pl, pr;
  uint32_t keep_q;
  float inv_keep;
  uint64_t seed;
  const uint64_t* epoch;
  const int32_t* lens;
  const uint8_t* drop_mask;
  float* zero_ptr;             // apply pass: a small buffer cleared for a LATER kernel (the layer's forward statistics)
  int zero_count;
  int rows_per_block;
  int red_raw;                 // apply pass: red[C:2C] holds raw sums of g*(z-mean) (from the GEMM epilogue, W2L_RED_RAW): scale by invstd here
};

// NaN passes through the activations, as torch.relu / torch.clamp do (fmaxf / fminf would swallow it): one instruction each
__device__ __forceinline__ float max_nan(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float min_nan(float a, float b) {
  float r;
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
template <int ACT>
__device__ __forceinline__ float act_fwd(float v) {
  if (ACT == W2L_ACT_RELU) return max_nan(v, 0.f);
  if (ACT == W2L_ACT_CLAMP20) return min_nan(max_nan(v, 0.f), 20.f);
  return v;
}
template <int ACT>
__device__ __forceinline__ bool act_pass(float pre) {
  if (ACT == W2L_ACT_RELU) return pre > 0.f;
  if (ACT == W2L_ACT_CLAMP20) return pre >= 0.f && pre <= 20.f;
  return true;
}

__device__ __forceinline__ void zero_small(float* p, int n) {
  if (p == nullptr || blockIdx.x != 0 || blockIdx.y != 0) return;
  for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < n; i += blockDim.x * blockDim.y) p[i] = 0.f;
}

// per-channel constants of a thread's 8 channels; with dropout the 1/keep factor is folded in, so that fma(z, sc, sh) is the value a
// KEPT element takes -- forward and backward build it with the same instructions, hence take the same activation-gate decisions
template <bool DROP, bool HAS_RES>
__device__ __forceinline__ void scale_for_dropout(float inv_keep, float (&sc)[8], float (&sh)[8], float (&rsc)[8], float (&rsh)[8]) {
  if (!DROP) return;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sc[i] *= inv_keep;
    sh[i] *= inv_keep;
    if (HAS_RES) {
      rsc[i] *= inv_keep;
      rsh[i] *= inv_keep;
    }
  }
}

template <typename TA, int ACT, bool DROP, bool HAS_RES>
__global__ void __launch_bounds__(kBnThreads, kBnFwdCtasPerSm) bn_act_pad_kernel(BnFwdArgs a) {
  const TA* az = reinterpret_cast<const TA*>(a.z);
  const TA* ares = reinterpret_cast<const TA*>(a.res);
  TA* ay = reinterpret_cast<TA*>(a.y);
  zero_small(a.zero_ptr, a.zero_count);
  const int ny = blockDim.y;
  const int cv = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = cv * 8;
  if (c >= a.C) return;
  float sc[8], sh[8], rsc[8], rsh[8];
  if (a.stats != nullptr) {                          // kernel-uniform: the finalize fold (the arithmetic of bn_finalize_kernel)
    float s1[8], s2[8], ga[8], be[8];
    load8f(a.stats + c, s1);
    load8f(a.stats + a.C + c, s2);
    if (a.gamma) load8f(a.gamma + c, ga);
    if (a.beta) load8f(a.beta + c, be);
    const bool writer = blockIdx.y == 0 && threadIdx.y == 0;
    if (writer && c == 0 && a.num_batches_tracked) *a.num_batches_tracked += 1;
    const double n = (double)a.stat_rows, inv_n = 1.0 / n;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const double mean = (double)s1[i] * inv_n;
      double var = (double)s2[i] * inv_n - mean * mean;          // the one place that needs the fp64 difference
      if (var < 0.0) var = 0.0;
      const float ve = (float)(var + (double)a.eps);
      float invstd = rsqrtf(ve);
      invstd = invstd * (1.5f - 0.5f * ve * invstd * invstd);     // one Newton step: <= 1 ulp
      const float g = a.gamma ? ga[i] : 1.f, bt = a.beta ? be[i] : 0.f;
      sc[i] = g * invstd;
      sh[i] = bt - (float)mean * g * invstd;
      if (writer) {
        a.fin[c + i] = sc[i];
        a.fin[a.C + c + i] = sh[i];
        a.fin[2 * a.C + c + i] = (float)mean;
        a.fin[3 * a.C + c + i] = invstd;
        if (a.running_mean) {
          const float m_full = (float)mean + (a.conv_bias ? a.conv_bias[c + i] : 0.f);
          a.running_mean[c + i] = (1.f - a.momentum) * a.running_mean[c + i] + a.momentum * m_full;
          const double unbiased = a.stat_rows > 1 ? var * n / (n - 1.0) : var;
          a.running_var[c + i] = (1.f - a.momentum) * a.running_var[c + i] + a.momentum * (float)unbiased;
        }
      }
    }
  } else {
    load8f(a.scale + c, sc);
    load8f(a.shift + c, sh);
  }
  if (HAS_RES) {
    load8f(a.res_scale + c, rsc);
    load8f(a.res_shift + c, rsh);
  }
  scale_for_dropout<DROP, HAS_RES>(a.inv_keep, sc, sh, rsc, rsh);
  const int rows = a.B * a.T, Tp = a.pl + a.T + a.pr;
  const int r_begin = blockIdx.y * a.rows_per_block, r_end = min(rows, r_begin + a.rows_per_block);   // r_begin % kRowGroup == 0
  constexpr int kBatch = (HAS_RES || sizeof(TA) == 4) ? 2 : kRowGroup;    // rows whose loads are in flight together (two streams with a residual)
  const uint64_t seed = DROP ? drop_seed(a) : 0;
  for (int r0 = r_begin + threadIdx.y * kRowGroup; r0 < r_end; r0 += ny * kRowGroup) {
    uint32_t keep = 0xFFFFFFFFu;
    if (DROP) keep = dropout_mask32(seed, (uint32_t)(r0 / kRowGroup), (uint32_t)cv, a.keep_q);
    int b = r0 / a.T, t = r0 - b * a.T;
#pragma unroll
    for (int h = 0; h < kRowGroup; h += kBatch) {
      Vec8<TA> zq[kBatch], rq[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        if (r0 + h + u < r_end) {
          const int64_t e = (int64_t)(r0 + h + u) * a.C + c;
          zq[u].load(az + e);
          if (HAS_RES) rq[u].load(ares + e);
        }
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int r = r0 + h + u;
        if (r < r_end) {
          float v[8];
          zq[u].unpack(v);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], sc[i], sh[i]);
          if (HAS_RES) {
            float rv[8];
            rq[u].unpack(rv);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += fmaf(rv[i], rsc[i], rsh[i]);
          }
          if (DROP && a.drop_mask) a.drop_mask[((int64_t)r * a.C + c) >> 3] = (uint8_t)(keep >> (8 * (h + u)));
          const bool masked = a.lens && t >= a.lens[b];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const bool on = !masked && (!DROP || ((keep >> (8 * (h + u) + i)) & 1u));
            v[i] = on ? act_fwd<ACT>(v[i]) : 0.f;
          }
          Vec8<TA> q;
          q.pack(v);
          TA* yb = ay + (int64_t)b * Tp * a.C + c;
          q.store(yb + (int64_t)(a.pl + t) * a.C);
          if (t >= 1 && t <= a.pl) q.store(yb + (int64_t)(a.pl - t) * a.C);                       // left mirror
          const int d = a.T - 1 - t;
          if (d >= 1 && d <= a.pr) q.store(yb + (int64_t)(a.pl + a.T - 1 + d) * a.C);              // right mirror
        }
        if (++t == a.T) {
          t = 0;
          ++b;
        }
      }
    }
  }
}

// one staged row of the backward passes: conv output, residual, upstream gradient, dropout keep-bits -- all requested together
template <typename TA>
struct BwdRow {
  Vec8<TA> z, r, d;
  uint32_t bits;
  int len;
};

template <typename TA, bool DROP, bool HAS_RES>
__device__ __forceinline__ void load_bwd_row(const BnBwdArgs& a, int r, int b, int t, int cv, BwdRow<TA>& in) {
  const int c = cv * 8;
  const int64_t e = (int64_t)r * a.C + c;
  in.z.load(reinterpret_cast<const TA*>(a.z) + e);
  if (HAS_RES) in.r.load(reinterpret_cast<const TA*>(a.res) + e);
  in.d.load(reinterpret_cast<const TA*>(a.dyp) + ((int64_t)b * (a.pl + a.T + a.pr) + a.pl + t) * a.C + c);
  // only LOADS here: a consumer of a load result between two rows' requests (the merge point of a "mask or hash" select was one)
  // makes the thread wait out the memory latency once per row instead of once per row group -- the regenerated-mask fallback
  // therefore lives in g_from_row
  in.bits = 0;
  if (DROP && a.drop_mask) in.bits = (uint32_t)__ldg(a.drop_mask + (e >> 3));
  in.len = a.lens ? __ldg(a.lens + b) : 0x7fffffff;
}

// masked upstream gradient g and the conv output for one staged row (reflect halo folded; activation gate, dropout, length mask)
template <typename TA, int ACT, bool DROP, bool HAS_RES>
__device__ __forceinline__ void g_from_row(const BnBwdArgs& a, int r, int b, int t, int c, const BwdRow<TA>& in, const float (&sc)[8],
                                           const float (&sh)[8], const float (&rsc)[8], const float (&rsh)[8], float (&g)[8],
                                           float (&zv)[8]) {
  in.z.unpack(zv);
  in.d.unpack(g);
  uint32_t bits = in.bits;
  if (DROP && !a.drop_mask)                                        // no stored keep-bits: regenerate them (kernel-uniform branch)
    bits = (dropout_mask32(drop_seed(a), (uint32_t)(r / kRowGroup), (uint32_t)(c >> 3), a.keep_q) >> (8 * (r % kRowGroup))) & 0xFFu;
  const int dr = a.T - 1 - t;
  if ((t >= 1 && t <= a.pl) || (dr >= 1 && dr <= a.pr)) {          // rows with a mirror image in the reflect halo (<8 % of the rows)
    const TA* base = reinterpret_cast<const TA*>(a.dyp) + (int64_t)b * (a.pl + a.T + a.pr) * a.C + c;
    float h[8];
    Vec8<TA> hq;
    if (t >= 1 && t <= a.pl) {
      hq.load(base + (int64_t)(a.pl - t) * a.C);
      hq.unpack(h);
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] += h[i];
    }
    if (dr >= 1 && dr <= a.pr) {
      hq.load(base + (int64_t)(a.pl + a.T - 1 + dr) * a.C);
      hq.unpack(h);
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] += h[i];
    }
  }
  float rv[8];
  if (HAS_RES) in.r.unpack(rv);
  const bool masked = t >= in.len;
  const float gs = DROP ? a.inv_keep : 1.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float pre = fmaf(zv[i], sc[i], sh[i]);                         // the kept element's value (1/keep folded into sc, sh)
    if (HAS_RES) pre += fmaf(rv[i], rsc[i], rsh[i]);
    const bool on = !masked && act_pass<ACT>(pre) && (!DROP || ((bits >> i) & 1u));
    g[i] = on ? g[i] * gs : 0.f;
  }
}

// red[0:C] += sum g, red[C:2C] += sum g*xhat   (accumulated as sum g*(z-mean), scaled by invstd once per CTA; red is zero on entry)
template <typename TA, int ACT, bool DROP, bool HAS_RES>
__global__ void __launch_bounds__(kBnThreads, kBnBwdCtasPerSm) bn_act_bwd_reduce_kernel(BnBwdArgs a) {
  constexpr int kRows = (HAS_RES || sizeof(TA) == 4) ? 2 : kRowGroup;     // a third stream / fp32 rows: fewer rows in flight, no spills
  __shared__ __align__(16) float s_a[kBnThreads * 8], s_b[kBnThreads * 8];
  const int bx = blockDim.x, ny = blockDim.y;
  const int cv = blockIdx.x * bx + threadIdx.x;
  const int c = cv * 8;
  float sg[8], sx[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) sg[i] = sx[i] = 0.f;
  if (c < a.C) {
    float sc[8], sh[8], rsc[8], rsh[8], mu[8];
    load8f(a.scale + c, sc);
    load8f(a.shift + c, sh);
    if (HAS_RES) {
      load8f(a.res_scale + c, rsc);
      load8f(a.res_shift + c, rsh);
    }
    scale_for_dropout<DROP, HAS_RES>(a.inv_keep, sc, sh, rsc, rsh);
    load8f(a.mean + c, mu);
    const int rows = a.B * a.T;
    const int r_begin = blockIdx.y * a.rows_per_block, r_end = min(rows, r_begin + a.rows_per_block);
    for (int r0 = r_begin + threadIdx.y * kRows; r0 < r_end; r0 += ny * kRows) {
      BwdRow<TA> in[kRows];
      const int b0 = r0 / a.T, t0 = r0 - b0 * a.T;
      {
        int b = b0, t = t0;
#pragma unroll
        for (int u = 0; u < kRows; ++u) {
          if (r0 + u < r_end) load_bwd_row<TA, DROP, HAS_RES>(a, r0 + u, b, t, cv, in[u]);
          if (++t == a.T) {
            t = 0;
            ++b;
          }
        }
      }
      int b = b0, t = t0;
#pragma unroll
      for (int u = 0; u < kRows; ++u) {
        if (r0 + u < r_end) {
          float g[8], zv[8];
          g_from_row<TA, ACT, DROP, HAS_RES>(a, r0 + u, b, t, c, in[u], sc, sh, rsc, rsh, g, zv);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            sg[i] += g[i];
            sx[i] = fmaf(g[i], zv[i] - mu[i], sx[i]);
          }
        }
        if (++t == a.T) {
          t = 0;
          ++b;
        }
      }
    }
    float is[8];
    load8f(a.invstd + c, is);
#pragma unroll
    for (int i = 0; i < 8; ++i) sx[i] *= is[i];
  }
  float* pa = s_a + (threadIdx.y * bx + threadIdx.x) * 8;
  float* pb = s_b + (threadIdx.y * bx + threadIdx.x) * 8;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    pa[i] = sg[i];
    pb[i] = sx[i];
  }
  __syncthreads();
  if (threadIdx.y == 0 && c < a.C) {                 // fold the row lanes in a fixed order, then ONE atomic per channel and CTA
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float u = 0.f, v = 0.f;
      for (int y = 0; y < ny; ++y) {
        u += s_a[(y * bx + threadIdx.x) * 8 + i];
        v += s_b[(y * bx + threadIdx.x) * 8 + i];
      }
      atomicAdd(a.red + c + i, u);
      atomicAdd(a.red + a.C + c + i, v);
    }
  }
}

// dz = gamma*invstd*(g - mean(g) - xhat*mean(g*xhat)) = A*g + Bz*z + Cc with per-channel A, Bz, Cc.
// dz [B, dz_rows, C]: rows [0, T) carry the gradient, rows [T, dz_rows) are zero-filled (the flat dgrad reads them as the zero
// padding between utterances).  Row ranges are taken in the REVERSE order of the reduce pass (last range first, each range from its
// end) so that the second read of (dy, z) starts with what the first pass touched last and is still in the 126 MB L2.
template <typename TA, int ACT, bool DROP, bool HAS_RES>
__global__ void __launch_bounds__(kBnThreads, kBnBwdCtasPerSm) bn_act_bwd_apply_kernel(BnBwdArgs a) {
  constexpr int kRows = (HAS_RES || sizeof(TA) == 4) ? 2 : kRowGroup;
  zero_small(a.zero_ptr, a.zero_count);
  const int ny = blockDim.y;
  const int cv = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = cv * 8;
  if (c >= a.C) return;
  const int rb = gridDim.y - 1 - blockIdx.y;
  float sc[8], sh[8], rsc[8], rsh[8];
  load8f(a.scale + c, sc);
  load8f(a.shift + c, sh);
  if (HAS_RES) {
    load8f(a.res_scale + c, rsc);
    load8f(a.res_shift + c, rsh);
  }
  scale_for_dropout<DROP, HAS_RES>(a.inv_keep, sc, sh, rsc, rsh);
  float kA[8], kB[8], kC[8];
  {
    float mu[8], is[8], sg[8], sx[8], ga[8];
    load8f(a.mean + c, mu);
    load8f(a.invstd + c, is);
    load8f(a.red + c, sg);
    load8f(a.red + a.C + c, sx);
    if (a.gamma) load8f(a.gamma + c, ga);
    if (a.red_raw) {
#pragma unroll
      for (int i = 0; i < 8; ++i) sx[i] *= is[i];
    }
    if (a.red_out && rb == 0 && threadIdx.y == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a.red_out[c + i] = sg[i];
        a.red_out[a.C + c + i] = sx[i];
      }
    }
    const float inv_m = 1.f / (float)((int64_t)a.B * a.T);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float coef = (a.gamma ? ga[i] : 1.f) * is[i];
      kA[i] = coef;
      kB[i] = -coef * sx[i] * inv_m * is[i];
      kC[i] = -coef * sg[i] * inv_m - kB[i] * mu[i];
    }
  }
  const int rows = a.B * a.T;
  const int r_begin = rb * a.rows_per_block, r_end = min(rows, r_begin + a.rows_per_block);
  const int groups = (r_end - r_begin + kRows - 1) / kRows;
  for (int gi = groups - 1 - (int)threadIdx.y; gi >= 0; gi -= ny) {
    const int r0 = r_begin + gi * kRows;
    BwdRow<TA> in[kRows];
    const int b0 = r0 / a.T, t0 = r0 - b0 * a.T;
    {
      int b = b0, t = t0;
#pragma unroll
      for (int u = 0; u < kRows; ++u) {
        if (r0 + u < r_end) load_bwd_row<TA, DROP, HAS_RES>(a, r0 + u, b, t, cv, in[u]);
        if (++t == a.T) {
          t = 0;
          ++b;
        }
      }
    }
    int b = b0, t = t0;
#pragma unroll
    for (int u = 0; u < kRows; ++u) {
      if (r0 + u < r_end) {
        float g[8], zv[8], o[8];
        g_from_row<TA, ACT, DROP, HAS_RES>(a, r0 + u, b, t, c, in[u], sc, sh, rsc, rsh, g, zv);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = fmaf(kA[i], g[i], fmaf(kB[i], zv[i], kC[i]));
        Vec8<TA> q;
        q.pack(o);
        q.store(reinterpret_cast<TA*>(a.dz) + ((int64_t)b * a.dz_rows + t) * a.C + c);
        if (a.g_out) {
          q.pack(g);
          q.store(reinterpret_cast<TA*>(a.g_out) + (int64_t)(r0 + u) * a.C + c);
        }
      }
      if (++t == a.T) {
        t = 0;
        ++b;
      }
    }
  }
  const int tail = a.dz_rows - a.T;
  if (tail > 0) {
    const int trows = a.B * tail;
    const int per = (trows + gridDim.y - 1) / gridDim.y;
    const int q_begin = blockIdx.y * per, q_end = min(trows, q_begin + per);
    for (int q = q_begin + threadIdx.y; q < q_end; q += ny) {
      const int b = q / tail, t = a.T + (q - b * tail);
      Vec8<TA> zv;
      zv.zero();
      zv.store(reinterpret_cast<TA*>(a.dz) + ((int64_t)b * a.dz_rows + t) * a.C + c);
    }
  }
}

// ---------------------------------------------------------------- log_softmax fwd / bwd (one warp per row)
__global__ void log_softmax_kernel(const float* __restrict__ logits, int ld, float* __restrict__ out, int64_t rows, int C, int mode,
                                   int32_t* __restrict__ nan_flag) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* src = logits + row * ld;
  float m = -INFINITY;
  for (int c = lane; c < C; c += 32) m = fmaxf(m, src[c]);
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += expf(src[c] - m);
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float lse = m + logf(s);
  bool bad = false;
  for (int c = lane; c < C; c += 32) {
    const float v = mode == 0 ? src[c] - lse : expf(src[c] - lse);
    bad |= v != v;
    out[row * C + c] = v;
  }
  // jasper.py:474 asserts that the scores hold no NaN: raise a flag here instead of a second full pass over them
  if (nan_flag && __any_sync(0xffffffffu, bad) && lane == 0) atomicOr(nan_flag, 1);
}

template <typename TOut>
__global__ void log_softmax_bwd_kernel(const float* __restrict__ g, const float* __restrict__ lp, const float* __restrict__ gscale,
                                       TOut* __restrict__ dlogits, int ld_out, int64_t rows, int C, int fused_identity) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float sc = gscale ? *gscale : 1.f;
  float s = 0.f;
  if (!fused_identity) {
    for (int c = lane; c < C; c += 32) s += g[row * C + c];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  for (int c = lane; c < ld_out; c += 32) {
    float v = 0.f;
    if (c < C) {
      v = g[row * C + c];
      if (!fused_identity) v -= expf(lp[row * C + c]) * s;
      v *= sc;
    }
    store_act(dlogits + row * ld_out + c, v);
  }
}

// column sums of a bf16 matrix (bias gradient); block (32, 8), 32 channels per block-column
template <typename TIn>
__global__ void colsum_kernel(const TIn* __restrict__ x, int64_t rows, int C, int ld, float* __restrict__ out,
                              int rows_per_block) {
  __shared__ float s[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int64_t r_begin = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r_end = min(rows, r_begin + rows_per_block);
  float acc = 0.f;
  if (c < C)
    for (int64_t r = r_begin + threadIdx.y; r < r_end; r += 8) acc += load_act(x + r * ld + c);
  s[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) t += s[y][threadIdx.x];
    atomicAdd(out + c, t);
  }
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  const int64_t n8 = n >> 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float v[8];
    load8f(src + i * 8, v);
    *reinterpret_cast<uint4*>(dst + i * 8) = pack8(v);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int64_t i = n8 * 8; i < n; ++i) dst[i] = __float2bfloat16_rn(src[i]);
}

// ---------------------------------------------------------------- Jasper length bookkeeping (jasper.py:91-95, 107-119)
// Every MaskedConv1d truncates the incoming lengths to integers (lens.to(long)), masks with them, and hands on
// (lens + 2p - d(k-1) - 1) / stride + 1 as a FLOAT tensor (true division); the reference spends ~6 tiny kernels per conv on
// this.  One thread per utterance walks the whole chain: out[j][b] = truncated length entering conv j (row n = the output
// lengths).  stride 0 marks a conv that does not mask (lengths pass through untouched).
constexpr int kMaxChain = 192;
struct LensChain {
  int32_t n;
  int16_t k[kMaxChain], s[kMaxChain], d[kMaxChain], p[kMaxChain];
};
__global__ void lens_chain_kernel(const void* __restrict__ lens_in, int is64, int B, LensChain c, int32_t* __restrict__ out,
                                  int64_t* __restrict__ final_out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  long long li = is64 ? reinterpret_cast<const long long*>(lens_in)[b] : (long long)reinterpret_cast<const int32_t*>(lens_in)[b];
  out[b] = (int32_t)li;
  for (int j = 0; j < c.n; ++j) {
    if (c.s[j] != 0) {
      const long long v = li + 2 * c.p[j] - c.d[j] * (c.k[j] - 1) - 1;
      const float lf = __fdiv_rn((float)v, (float)c.s[j]) + 1.f;       // int64 tensor / int -> float32 true division, then + 1
      li = (long long)lf;                                               // the next conv's lens.to(dtype=torch.long)
    }
    out[(int64_t)(j + 1) * B + b] = (int32_t)li;
  }
  if (final_out) final_out[b] = li;
}

static inline int grid_for(int64_t items, int threads) {
  int64_t blocks = (items + threads - 1) / threads;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// launch geometry of the three BatchNorm / activation passes (see the comment above their kernels)
struct BnGeo {
  dim3 grid, block;
  int rows_per_block;
};
static BnGeo bn_geo(int64_t rows, int C, int ctas_per_sm) {
  BnGeo g;
  const int cvecs = C / 8;
  const int bx = cvecs < kBnThreads ? cvecs : kBnThreads;
  const int ny = kBnThreads / bx;
  const int col_blocks = (cvecs + bx - 1) / bx;
  int64_t target = (int64_t)num_sms() * ctas_per_sm / col_blocks;
  if (target < 1) target = 1;
  const int64_t quantum = (int64_t)ny * kRowGroup;                // whole row groups per row lane
  int64_t rpb = (rows + target - 1) / target;
  rpb = (rpb + quantum - 1) / quantum * quantum;
  g.rows_per_block = (int)rpb;
  g.grid = dim3((unsigned)col_blocks, (unsigned)((rows + rpb - 1) / rpb));
  g.block = dim3((unsigned)bx, (unsigned)ny);
  return g;
}

// compile-time (activation, dropout, residual) variants of the kernels above
#define W2L_BN_DISPATCH_T(KERNEL, TA, act, drop, has_res, ...)                                                          \
  do {                                                                                                                 \
    const int key_ = (act) * 4 + ((drop) ? 2 : 0) + ((has_res) ? 1 : 0);                                               \
    switch (key_) {                                                                                                    \
      case 0: KERNEL<TA, W2L_ACT_NONE, false, false><<<geo.grid, geo.block, 0, st>>>(__VA_ARGS__); break;              \
      case 1: KERNEL<TA, W2L_ACT_NONE, false, true><<<geo.grid, geo.block, 0, st>>>(__VA_ARGS__); break;               \
      case 2: KERNEL<TA, W2L_ACT_NONE, true, false><<<geo.grid, geo.block, 0, st>>>(__VA_ARGS__); break;               \
      case 3: KERNEL<TA, W2L_ACT_NONE, true, true><<<geo.grid, geo.block, 0, st>>>(__VA_ARGS__); break;                \
      case 4: KERNEL<TA, W2L_ACT_RELU, false, false><<<geo.grid, geo.block, 0, st>>>(__VA_ARGS__); break;              \
      case 5: KERNEL<TA, W2L_ACT_RELU, false, true><<<geo.grid, geo.block, 0, st>>>(__VA_ARGS__); break;               \
      case 6: KERNEL<TA, W2L_ACT_RELU, true, false><<<geo.grid, geo.block, 0, st>>>(__VA_ARGS__); break;               \
      case 7: KERNEL<TA, W2L_ACT_RELU, true, true><<<geo.grid, geo.block, 0, st>>>(__VA_ARGS__); break;                \
      case 8: KERNEL<TA, W2L_ACT_CLAMP20, false, false><<<geo.grid, geo.block, 0, st>>>(__VA_ARGS__); break;           \
      case 9: KERNEL<TA, W2L_ACT_CLAMP20, false, true><<<geo.grid, geo.block, 0, st>>>(__VA_ARGS__); break;            \
      case 10: KERNEL<TA, W2L_ACT_CLAMP20, true, false><<<geo.grid, geo.block, 0, st>>>(__VA_ARGS__); break;           \
      default: KERNEL<TA, W2L_ACT_CLAMP20, true, true><<<geo.grid, geo.block, 0, st>>>(__VA_ARGS__); break;            \
    }                                                                                                                  \
  } while (0)
// `act` may carry W2L_STORE_F32: the activation buffers of the call are fp32 (kernels instantiated with TA = float)
#define W2L_BN_DISPATCH(KERNEL, act, drop, has_res, ...)                                                                \
  do {                                                                                                                 \
    if ((act) & W2L_STORE_F32)                                                                                         \
      W2L_BN_DISPATCH_T(KERNEL, float, (act) & 0xFF, drop, has_res, __VA_ARGS__);                                      \
    else                                                                                                               \
      W2L_BN_DISPATCH_T(KERNEL, __nv_bfloat16, (act) & 0xFF, drop, has_res, __VA_ARGS__);                              \
  } while (0)

// dropout probability -> (keep_q, inv_keep) of the kernels' 12-bit keep probability; every pass of a layer derives them the same way
static void drop_quant(float drop_p, uint32_t* keep_q, float* inv_keep) {
  if (!(drop_p > 0.f)) {
    *keep_q = 0;
    *inv_keep = 1.f;
    return;
  }
  const uint32_t one = 1u << kDropBits;
  uint32_t q = (uint32_t)((1.0 - (double)drop_p) * (double)one + 0.5);
  if (q < 1) q = 1;
  if (q > one - 1) q = one - 1;
  *keep_q = q;
  *inv_keep = (float)((double)one / (double)q);
}

static int check_bn_args(const char* who, const void* z, const void* res, const float* res_scale, const float* res_shift, int B, int T,
                         int C, int pl, int pr, int act, float drop_p) {
  W2L_REQUIRE(z != nullptr, "%s: null pointer", who);
  W2L_REQUIRE(!res || (res_scale && res_shift), "%s: residual needs res_scale/res_shift", who);
  W2L_REQUIRE(B >= 1 && T >= 1 && C >= 8 && C % 8 == 0, "%s: bad shape B=%d T=%d C=%d (C must be a multiple of 8)", who, B, T, C);
  W2L_REQUIRE(pl >= 0 && pr >= 0 && pl < T && pr < T, "%s: reflect halo (%d,%d) must be smaller than T=%d", who, pl, pr, T);
  W2L_REQUIRE((int64_t)B * (T + pl + pr) < (1ll << 31) / 8, "%s: B*T too large", who);
  W2L_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "%s: dropout p=%f out of [0,1)", who, drop_p);
  W2L_REQUIRE((act & ~(W2L_STORE_F32 | W2L_RED_RAW)) >= 0 && (act & ~(W2L_STORE_F32 | W2L_RED_RAW)) <= 2, "%s: unknown activation %d", who, act);
  return W2L_OK;
}

static int launch_bn_fwd(BnFwdArgs& a, int act, const void* res, void* stream) {
  const BnGeo geo = bn_geo((int64_t)a.B * a.T, a.C, kBnFwdCtasPerSm);
  a.rows_per_block = geo.rows_per_block;
  cudaStream_t st = (cudaStream_t)stream;
  W2L_BN_DISPATCH(bn_act_pad_kernel, act, a.keep_q != 0, res != nullptr, a);
  return after_launch("bn_act_pad_kernel");
}

static int rows_per_block_for(int64_t rows, int col_blocks) {
  int64_t target_blocks = (int64_t)num_sms() * 8 / (col_blocks > 0 ? col_blocks : 1);
  if (target_blocks < 1) target_blocks = 1;
  int64_t rpb = (rows + target_blocks - 1) / target_blocks;
  if (rpb < 32) rpb = 32;
  return (int)rpb;
}

}  // namespace w2l

extern "C" {

int w2l_set_dropout_epoch(const uint64_t* epoch) {
  w2l::g_dropout_epoch = epoch;
  return W2L_OK;
}

static int im2col_ncw_impl(int f32, const float* x, void* out, int32_t B, int32_t F, int32_t T, int32_t rows, int32_t k, int32_t stride,
                           int32_t dilation, int32_t pad_left, int32_t pad_mode, const int32_t* lens, void* stream) {
  using namespace w2l;
  W2L_REQUIRE(x && out, "im2col_ncw: null pointer");
  W2L_REQUIRE(B >= 1 && F >= 1 && T >= 1 && rows >= 1 && k >= 1 && stride >= 1 && dilation >= 1 && pad_left >= 0, "im2col_ncw: bad geometry");
  W2L_REQUIRE(pad_mode != W2L_PAD_REFLECT || pad_left < T, "im2col_ncw: reflect pad %d must be < T=%d", pad_left, T);
  W2L_REQUIRE(B <= 65535, "im2col_ncw: B too large");
  const int span = 31 * stride + (k - 1) * dilation + 1;
  const int pitch = span | 1;
  const size_t smem = (size_t)F * pitch * sizeof(float);
  W2L_REQUIRE(smem <= 200 * 1024, "im2col_ncw: F=%d k=%d needs %zu bytes of shared memory", F, k, smem);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    W2L_CUDA(cudaFuncSetAttribute(im2col_ncw_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    W2L_CUDA(cudaFuncSetAttribute(im2col_ncw_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  dim3 grid((rows + 31) / 32, B);
  if (f32)
    im2col_ncw_kernel<float><<<grid, 256, smem, (cudaStream_t)stream>>>(x, (float*)out, F, T, rows, k, stride, dilation, pad_left, pad_mode,
                                                                        lens, span, pitch);
  else
    im2col_ncw_kernel<__nv_bfloat16><<<grid, 256, smem, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)out, F, T, rows, k, stride, dilation,
                                                                                pad_left, pad_mode, lens, span, pitch);
  return after_launch("im2col_ncw_kernel");
}
int w2l_im2col_ncw(const float* x, void* out, int32_t B, int32_t F, int32_t T, int32_t rows, int32_t k, int32_t stride,
                   int32_t dilation, int32_t pad_left, int32_t pad_mode, const int32_t* lens, void* stream) {
  return im2col_ncw_impl(0, x, out, B, F, T, rows, k, stride, dilation, pad_left, pad_mode, lens, stream);
}
int w2l_im2col_ncw_f32(const float* x, float* out, int32_t B, int32_t F, int32_t T, int32_t rows, int32_t k, int32_t stride,
                       int32_t dilation, int32_t pad_left, int32_t pad_mode, const int32_t* lens, void* stream) {
  return im2col_ncw_impl(1, x, out, B, F, T, rows, k, stride, dilation, pad_left, pad_mode, lens, stream);
}

int w2l_im2col_tm(const void* x, void* out, int32_t B, int32_t x_rows, int32_t C, int32_t T_out, int32_t k, int32_t stride,
                  int32_t dilation, int32_t pad_left, void* stream) {
  using namespace w2l;
  W2L_REQUIRE(x && out, "im2col_tm: null pointer");
  W2L_REQUIRE(B >= 1 && x_rows >= 1 && T_out >= 1 && k >= 1 && stride >= 1 && dilation >= 1 && pad_left >= 0, "im2col_tm: bad geometry");
  W2L_REQUIRE(C >= 8 && C % 8 == 0, "im2col_tm: C=%d must be a multiple of 8", C);
  W2L_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 15) == 0, "im2col_tm: pointers must be 16-byte aligned");
  const int64_t total = (int64_t)B * T_out * k * (C / 8);
  im2col_tm_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, B, x_rows, C,
                                                                         T_out, k, stride, dilation, pad_left);
  return after_launch("im2col_tm_kernel");
}

int w2l_col2im_tm(const void* dcol, void* dx, int32_t B, int32_t x_rows, int32_t C, int32_t T_out, int32_t k, int32_t stride,
                  int32_t dilation, int32_t pad_left, void* stream) {
  using namespace w2l;
  W2L_REQUIRE(dcol && dx, "col2im_tm: null pointer");
  W2L_REQUIRE(B >= 1 && x_rows >= 1 && T_out >= 1 && k >= 1 && stride >= 1 && dilation >= 1 && pad_left >= 0, "col2im_tm: bad geometry");
  W2L_REQUIRE(C >= 8 && C % 8 == 0, "col2im_tm: C=%d must be a multiple of 8", C);
  W2L_REQUIRE(((uintptr_t)dcol & 15) == 0 && ((uintptr_t)dx & 15) == 0, "col2im_tm: pointers must be 16-byte aligned");
  const int64_t total = (int64_t)B * x_rows * (C / 8);
  col2im_tm_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dcol, (__nv_bfloat16*)dx, B, x_rows, C,
                                                                         T_out, k, stride, dilation, pad_left);
  return after_launch("col2im_tm_kernel");
}

int w2l_ncw_to_tm(const float* x, void* out, int32_t B, int32_t C, int32_t T, void* stream) {
  return w2l_im2col_ncw(x, out, B, C, T, T, 1, 1, 1, 0, W2L_PAD_ZERO, nullptr, stream);
}

int w2l_tm_to_ncw(const void* x, int32_t x_dtype, float* out, int32_t B, int32_t T, int32_t C, int32_t x_rows, int32_t x_row_offset,
                  int32_t ld, void* stream) {
  using namespace w2l;
  W2L_REQUIRE(x && out && B >= 1 && T >= 1 && C >= 1 && ld >= C && x_rows >= T + x_row_offset, "tm_to_ncw: bad arguments");
  W2L_REQUIRE(B <= 65535 && (C + 31) / 32 <= 65535, "tm_to_ncw: shape too large");
  dim3 grid((T + 31) / 32, (C + 31) / 32, B), block(32, 8);
  const int64_t bs = (int64_t)x_rows * ld;
  if (x_dtype == W2L_DTYPE_BF16)
    tm_to_ncw_kernel<__nv_bfloat16><<<grid, block, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, out, T, C, bs, x_row_offset, ld, T);
  else
    tm_to_ncw_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>((const float*)x, out, T, C, bs, x_row_offset, ld, T);
  return after_launch("tm_to_ncw_kernel");
}

int w2l_tm_to_ct_f32(const float* x, float* out, int32_t B, int32_t T, int32_t C, int32_t x_rows, int32_t x_row_offset, int32_t ld,
                     int32_t out_pitch, void* stream) {
  using namespace w2l;
  W2L_REQUIRE(x && out && B >= 1 && T >= 1 && C >= 1 && ld >= C && x_rows >= T + x_row_offset && out_pitch >= T, "tm_to_ct_f32: bad arguments");
  W2L_REQUIRE(B <= 65535 && (C + 31) / 32 <= 65535, "tm_to_ct_f32: shape too large");
  dim3 grid((T + 31) / 32, (C + 31) / 32, B), block(32, 8);
  tm_to_ncw_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>(x, out, T, C, (int64_t)x_rows * ld, x_row_offset, ld, out_pitch);
  return after_launch("tm_to_ncw_kernel");
}

int w2l_bn_stats(const void* z, int64_t rows, int32_t C, float* stats, void* stream) {
  using namespace w2l;
  W2L_REQUIRE(z && stats && rows >= 1 && C >= 8 && C % 8 == 0, "bn_stats: bad arguments (C must be a multiple of 8)");
  const int col_blocks = (C + 255) / 256;
  const int rpb = rows_per_block_for(rows, col_blocks);
  dim3 grid(col_blocks, (unsigned)((rows + rpb - 1) / rpb)), block(32, 8);
  bn_stats_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)z, rows, C, stats, rpb);
  return after_launch("bn_stats_kernel");
}

int w2l_bn_finalize(const float* stats, int64_t rows, int32_t C, const float* gamma, const float* beta, const float* conv_bias,
                    float eps, float momentum, float* running_mean, float* running_var, float* scale, float* shift, float* mean,
                    float* invstd, int64_t* num_batches_tracked, void* stream) {
  using namespace w2l;
  W2L_REQUIRE(stats && scale && shift && rows >= 1 && C >= 1, "bn_finalize: bad arguments");
  W2L_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "bn_finalize: running stats must be given together");
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(stats, rows, C, gamma, beta, conv_bias, eps, momentum,
                                                                        running_mean, running_var, scale, shift, mean, invstd,
                                                                        num_batches_tracked);
  return after_launch("bn_finalize_kernel");
}

int w2l_bn_act_pad(const void* z, const float* scale, const float* shift, const void* res, const float* res_scale,
                   const float* res_shift, void* y, int32_t B, int32_t T, int32_t C, int32_t pad_left, int32_t pad_right,
                   int32_t act, float drop_p, uint64_t seed, const int32_t* lens, void* drop_mask, void* stream) {
  using namespace w2l;
  int rc = check_bn_args("bn_act_pad", z, res, res_scale, res_shift, B, T, C, pad_left, pad_right, act, drop_p);
  if (rc) return rc;
  W2L_REQUIRE(y && scale && shift, "bn_act_pad: null pointer");
  BnFwdArgs a;
  memset(&a, 0, sizeof(a));
  a.z = z;
  a.res = res;
  a.scale = scale;
  a.shift = shift;
  a.res_scale = res_scale;
  a.res_shift = res_shift;
  a.y = y;
  a.B = B, a.T = T, a.C = C, a.pl = pad_left, a.pr = pad_right;
  drop_quant(drop_p, &a.keep_q, &a.inv_keep);
  a.seed = seed;
  a.epoch = g_dropout_epoch;
  a.lens = lens;
  a.drop_mask = (uint8_t*)drop_mask;
  return launch_bn_fwd(a, act, res, stream);
}

int w2l_bn_finalize_act_pad(const void* z, const float* stats, int64_t stat_rows, const float* gamma, const float* beta,
                            const float* conv_bias, float eps, float momentum, float* running_mean, float* running_var,
                            int64_t* num_batches_tracked, float* fin, const void* res, const float* res_scale, const float* res_shift,
                            void* y, int32_t B, int32_t T, int32_t C, int32_t pad_left, int32_t pad_right, int32_t act, float drop_p,
                            uint64_t seed, const int32_t* lens, void* drop_mask, float* zero_ptr, int32_t zero_count, void* stream) {
  using namespace w2l;
  int rc = check_bn_args("bn_finalize_act_pad", z, res, res_scale, res_shift, B, T, C, pad_left, pad_right, act, drop_p);
  if (rc) return rc;
  W2L_REQUIRE(y && stats && fin && stat_rows >= 1, "bn_finalize_act_pad: null pointer / no rows");
  W2L_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "bn_finalize_act_pad: running stats must be given together");
  W2L_REQUIRE(zero_count >= 0 && (zero_ptr != nullptr || zero_count == 0), "bn_finalize_act_pad: bad zero buffer");
  BnFwdArgs a;
  memset(&a, 0, sizeof(a));
  a.z = z;
  a.res = res;
  a.res_scale = res_scale;
  a.res_shift = res_shift;
  a.stats = stats;
  a.stat_rows = stat_rows;
  a.gamma = gamma;
  a.beta = beta;
  a.conv_bias = conv_bias;
  a.eps = eps;
  a.momentum = momentum;
  a.running_mean = running_mean;
  a.running_var = running_var;
  a.num_batches_tracked = num_batches_tracked;
  a.fin = fin;
  a.y = y;
  a.B = B, a.T = T, a.C = C, a.pl = pad_left, a.pr = pad_right;
  drop_quant(drop_p, &a.keep_q, &a.inv_keep);
  a.seed = seed;
  a.epoch = g_dropout_epoch;
  a.lens = lens;
  a.drop_mask = (uint8_t*)drop_mask;
  a.zero_ptr = zero_ptr;
  a.zero_count = zero_count;
  return launch_bn_fwd(a, act, res, stream);
}

static int fill_bwd_args(w2l::BnBwdArgs& a, const char* who, const void* dyp, const void* z, const void* res, const float* scale,
                         const float* shift, const float* res_scale, const float* res_shift, const float* mean, const float* invstd,
                         float* red, int32_t B, int32_t T, int32_t C, int32_t pad_left, int32_t pad_right, int32_t act, float drop_p,
                         uint64_t seed, const int32_t* lens, const void* drop_mask) {
  using namespace w2l;
  int rc = check_bn_args(who, z, res, res_scale, res_shift, B, T, C, pad_left, pad_right, act, drop_p);
  if (rc) return rc;
  W2L_REQUIRE(dyp && scale && shift && mean && invstd && red, "%s: null pointer", who);
  memset(&a, 0, sizeof(a));
  a.z = z;
  a.res = res;
  a.dyp = dyp;
  a.scale = scale;
  a.shift = shift;
  a.res_scale = res_scale;
  a.res_shift = res_shift;
  a.mean = mean;
  a.invstd = invstd;
  a.red = red;
  a.B = B, a.T = T, a.C = C, a.pl = pad_left, a.pr = pad_right;
  drop_quant(drop_p, &a.keep_q, &a.inv_keep);
  a.seed = seed;
  a.epoch = g_dropout_epoch;
  a.lens = lens;
  a.drop_mask = (const uint8_t*)drop_mask;
  return W2L_OK;
}

int w2l_bn_act_bwd_reduce(const void* dyp, const void* z, const void* res, const float* scale, const float* shift,
                          const float* res_scale, const float* res_shift, const float* mean, const float* invstd, float* red,
                          int32_t B, int32_t T, int32_t C, int32_t pad_left, int32_t pad_right, int32_t act, float drop_p,
                          uint64_t seed, const int32_t* lens, const void* drop_mask, void* stream) {
  using namespace w2l;
  BnBwdArgs a;
  int rc = fill_bwd_args(a, "bn_act_bwd_reduce", dyp, z, res, scale, shift, res_scale, res_shift, mean, invstd, red, B, T, C, pad_left,
                         pad_right, act, drop_p, seed, lens, drop_mask);
  if (rc) return rc;
  const BnGeo geo = bn_geo((int64_t)B * T, C, kBnBwdCtasPerSm);
  a.rows_per_block = geo.rows_per_block;
  cudaStream_t st = (cudaStream_t)stream;
  W2L_BN_DISPATCH(bn_act_bwd_reduce_kernel, act, a.keep_q != 0, res != nullptr, a);
  return after_launch("bn_act_bwd_reduce_kernel");
}

int w2l_bn_act_bwd_apply(const void* dyp, const void* z, const void* res, const float* scale, const float* shift,
                         const float* res_scale, const float* res_shift, const float* mean, const float* invstd, const float* gamma,
                         const float* red, void* dz, int32_t dz_rows, void* g_out, int32_t B, int32_t T, int32_t C,
                         int32_t pad_left, int32_t pad_right, int32_t act, float drop_p, uint64_t seed, const int32_t* lens,
                         const void* drop_mask, float* red_out, float* zero_ptr, int32_t zero_count, void* stream) {
  using namespace w2l;
  BnBwdArgs a;
  int rc = fill_bwd_args(a, "bn_act_bwd_apply", dyp, z, res, scale, shift, res_scale, res_shift, mean, invstd, const_cast<float*>(red), B, T,
                         C, pad_left, pad_right, act, drop_p, seed, lens, drop_mask);
  if (rc) return rc;
  W2L_REQUIRE(dz != nullptr && dz_rows >= T, "bn_act_bwd_apply: null dz or dz_rows %d < T %d", dz_rows, T);
  W2L_REQUIRE(zero_count >= 0 && (zero_ptr != nullptr || zero_count == 0), "bn_act_bwd_apply: bad zero buffer");
  W2L_REQUIRE(zero_ptr == nullptr || zero_ptr != red, "bn_act_bwd_apply: the buffer to clear must not be the reduction this pass reads");
  a.gamma = gamma;
  a.red_raw = (act & W2L_RED_RAW) != 0;
  a.dz = dz;
  a.dz_rows = dz_rows;
  a.g_out = g_out;
  a.red_out = red_out;
  a.zero_ptr = zero_ptr;
  a.zero_count = zero_count;
  const BnGeo geo = bn_geo((int64_t)B * T, C, kBnBwdCtasPerSm);
  a.rows_per_block = geo.rows_per_block;
  cudaStream_t st = (cudaStream_t)stream;
  W2L_BN_DISPATCH(bn_act_bwd_apply_kernel, act, a.keep_q != 0, res != nullptr, a);
  return after_launch("bn_act_bwd_apply_kernel");
}

int w2l_log_softmax(const float* logits, int32_t ld, float* out, int64_t rows, int32_t C, int32_t mode, int32_t* nan_flag, void* stream) {
  using namespace w2l;
  W2L_REQUIRE(logits && out && rows >= 1 && C >= 1 && ld >= C, "log_softmax: bad arguments");
  log_softmax_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(logits, ld, out, rows, C, mode, nan_flag);
  return after_launch("log_softmax_kernel");
}

static int log_softmax_bwd_impl(int f32, const float* g, const float* lp, const float* gscale, void* dlogits, int32_t ld_out, int64_t rows,
                                int32_t C, int32_t fused_identity, void* stream) {
  using namespace w2l;
  W2L_REQUIRE(g && dlogits && rows >= 1 && C >= 1 && ld_out >= C, "log_softmax_bwd: bad arguments");
  W2L_REQUIRE(fused_identity || lp, "log_softmax_bwd: log-probs required");
  if (f32)
    log_softmax_bwd_kernel<float><<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(g, lp, gscale, (float*)dlogits, ld_out, rows, C,
                                                                                               fused_identity);
  else
    log_softmax_bwd_kernel<__nv_bfloat16><<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(g, lp, gscale, (__nv_bfloat16*)dlogits,
                                                                                                       ld_out, rows, C, fused_identity);
  return after_launch("log_softmax_bwd_kernel");
}
int w2l_log_softmax_bwd(const float* g, const float* lp, const float* gscale, void* dlogits, int32_t ld_out, int64_t rows,
                        int32_t C, int32_t fused_identity, void* stream) {
  return log_softmax_bwd_impl(0, g, lp, gscale, dlogits, ld_out, rows, C, fused_identity, stream);
}
int w2l_log_softmax_bwd_f32(const float* g, const float* lp, const float* gscale, float* dlogits, int32_t ld_out, int64_t rows,
                            int32_t C, int32_t fused_identity, void* stream) {
  return log_softmax_bwd_impl(1, g, lp, gscale, dlogits, ld_out, rows, C, fused_identity, stream);
}

static int colsum_impl(int f32, const void* x, int64_t rows, int32_t C, int32_t ld, float* out, void* stream) {
  using namespace w2l;
  W2L_REQUIRE(x && out && rows >= 1 && C >= 1 && ld >= C, "colsum: bad arguments");
  const int col_blocks = (C + 31) / 32;
  const int rpb = rows_per_block_for(rows, col_blocks);
  dim3 grid(col_blocks, (unsigned)((rows + rpb - 1) / rpb)), block(32, 8);
  if (f32)
    colsum_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>((const float*)x, rows, C, ld, out, rpb);
  else
    colsum_kernel<__nv_bfloat16><<<grid, block, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, rows, C, ld, out, rpb);
  return after_launch("colsum_kernel");
}
int w2l_colsum(const void* x, int64_t rows, int32_t C, int32_t ld, float* out, void* stream) { return colsum_impl(0, x, rows, C, ld, out, stream); }
int w2l_colsum_f32(const float* x, int64_t rows, int32_t C, int32_t ld, float* out, void* stream) {
  return colsum_impl(1, x, rows, C, ld, out, stream);
}

int w2l_cast_bf16(const float* src, void* dst, int64_t n, void* stream) {
  using namespace w2l;
  W2L_REQUIRE(src && dst && n >= 0, "cast_bf16: bad arguments");
  if (n == 0) return W2L_OK;
  W2L_REQUIRE(((uintptr_t)src & 31) == 0 && ((uintptr_t)dst & 15) == 0, "cast_bf16: pointers must be 32/16-byte aligned");
  cast_bf16_kernel<<<grid_for(n / 8 + 1, 256), 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, n);
  return after_launch("cast_bf16_kernel");
}

}  // extern "C"

// ---------------------------------------------------------------- reflect halo fill (inference path)
// y [B, pl+T+pr, C] bf16 whose interior rows [pl, pl+T) are already written: fills the mirrored halo rows in place.
namespace w2l {
template <typename TA>
__global__ void reflect_halo_kernel(TA* __restrict__ y, int B, int T, int C, int pl, int pr) {
  const int c8 = C >> 3, halo = pl + pr, Tp = pl + T + pr;
  const int64_t total = (int64_t)B * halo * c8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c8) * 8;
    const int64_t bh = i / c8;
    const int h = (int)(bh % halo), b = (int)(bh / halo);
    // halo row index and its mirror source (both in padded coordinates)
    const int dst = h < pl ? h : pl + T + (h - pl);
    const int src = h < pl ? 2 * pl - h : pl + T - 2 - (h - pl);
    TA* yb = y + (int64_t)b * Tp * C + c;
    Vec8<TA> q;
    q.load(yb + (int64_t)src * C);
    q.store(yb + (int64_t)dst * C);
  }
}
}  // namespace w2l

namespace w2l {
static int reflect_halo_impl(int f32, void* y, int32_t B, int32_t T, int32_t C, int32_t pad_left, int32_t pad_right, void* stream) {
  using namespace w2l;
  W2L_REQUIRE(y && B >= 1 && T >= 1 && C >= 8 && C % 8 == 0, "reflect_halo: bad arguments");
  W2L_REQUIRE(pad_left >= 0 && pad_right >= 0 && pad_left < T && pad_right < T, "reflect_halo: halo must be smaller than T");
  if (pad_left + pad_right == 0) return W2L_OK;
  const int64_t total = (int64_t)B * (pad_left + pad_right) * (C / 8);
  if (f32)
    reflect_halo_kernel<float><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((float*)y, B, T, C, pad_left, pad_right);
  else
    reflect_halo_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)y, B, T, C, pad_left, pad_right);
  return after_launch("reflect_halo_kernel");
}
}  // namespace w2l
extern "C" int w2l_reflect_halo(void* y, int32_t B, int32_t T, int32_t C, int32_t pad_left, int32_t pad_right, void* stream) {
  return w2l::reflect_halo_impl(0, y, B, T, C, pad_left, pad_right, stream);
}
extern "C" int w2l_reflect_halo_f32(float* y, int32_t B, int32_t T, int32_t C, int32_t pad_left, int32_t pad_right, void* stream) {
  return w2l::reflect_halo_impl(1, y, B, T, C, pad_left, pad_right, stream);
}

// ---------------------------------------------------------------- transposed weight shadow for backward-data
// w [k, Co, Ci] fp32 (kernel-layout master weights) -> wt [k, Ci_pad, Co_pad] bf16 with wt[k-1-j][ci][co] = w[j][co][ci]
namespace w2l {
template <typename TOut>
__global__ void pack_wt_kernel(const float* __restrict__ w, TOut* __restrict__ wt, int k, int Co, int Ci, int Co_pad, int Ci_pad) {
  __shared__ float tile[32][33];
  const int j = blockIdx.z, co0 = blockIdx.y * 32, ci0 = blockIdx.x * 32;
  const float* src = w + (int64_t)j * Co * Ci;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int co = co0 + i, ci = ci0 + threadIdx.x;
    tile[i][threadIdx.x] = (co < Co && ci < Ci) ? src[(int64_t)co * Ci + ci] : 0.f;
  }
  __syncthreads();
  TOut* dst = wt + (int64_t)(k - 1 - j) * Ci_pad * Co_pad;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int ci = ci0 + i, co = co0 + threadIdx.x;
    if (ci < Ci_pad && co < Co_pad) store_act(dst + (int64_t)ci * Co_pad + co, tile[threadIdx.x][i]);
  }
}
}  // namespace w2l

namespace w2l {
static int pack_wt_impl(int f32, const float* w, void* wt, int32_t k, int32_t Cout, int32_t Cin, int32_t Cout_pad, int32_t Cin_pad, void* stream) {
  using namespace w2l;
  W2L_REQUIRE(w && wt && k >= 1 && Cout >= 1 && Cin >= 1 && Cout_pad >= Cout && Cin_pad >= Cin, "pack_wt: bad arguments");
  W2L_REQUIRE(k <= 65535 && (Cout_pad + 31) / 32 <= 65535, "pack_wt: shape too large");
  dim3 grid((Cin_pad + 31) / 32, (Cout_pad + 31) / 32, k), block(32, 8);
  if (f32)
    pack_wt_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>(w, (float*)wt, k, Cout, Cin, Cout_pad, Cin_pad);
  else
    pack_wt_kernel<__nv_bfloat16><<<grid, block, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)wt, k, Cout, Cin, Cout_pad, Cin_pad);
  return after_launch("pack_wt_kernel");
}
}  // namespace w2l
extern "C" int w2l_pack_wt(const float* w, void* wt, int32_t k, int32_t Cout, int32_t Cin, int32_t Cout_pad, int32_t Cin_pad, void* stream) {
  return w2l::pack_wt_impl(0, w, wt, k, Cout, Cin, Cout_pad, Cin_pad, stream);
}
extern "C" int w2l_pack_wt_f32(const float* w, float* wt, int32_t k, int32_t Cout, int32_t Cin, int32_t Cout_pad, int32_t Cin_pad, void* stream) {
  return w2l::pack_wt_impl(1, w, wt, k, Cout, Cin, Cout_pad, Cin_pad, stream);
}

extern "C" int w2l_lens_chain(const void* lens_in, int32_t lens_is_int64, int32_t B, const int32_t* conv_params_host, int32_t n_convs,
                              int32_t* lens_out, int64_t* final_out, void* stream) {
  using namespace w2l;
  W2L_REQUIRE(lens_in && lens_out && B >= 1 && n_convs >= 0 && n_convs <= kMaxChain, "lens_chain: bad arguments (at most %d convs)", kMaxChain);
  W2L_REQUIRE(n_convs == 0 || conv_params_host, "lens_chain: null conv table");
  LensChain c;
  c.n = n_convs;
  for (int j = 0; j < n_convs; ++j) {
    const int32_t* q = conv_params_host + 4 * j;            // (kernel, stride [0 = unmasked pass-through], dilation, padding)
    W2L_REQUIRE(q[0] >= 1 && q[0] < 32768 && q[1] >= 0 && q[1] < 32768 && q[2] >= 1 && q[2] < 32768 && q[3] >= 0 && q[3] < 32768,
                "lens_chain: conv %d has out-of-range geometry", j);
    c.k[j] = (int16_t)q[0];
    c.s[j] = (int16_t)q[1];
    c.d[j] = (int16_t)q[2];
    c.p[j] = (int16_t)q[3];
  }
  lens_chain_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(lens_in, lens_is_int64, B, c, lens_out, final_out);
  return after_launch("lens_chain_kernel");
}
