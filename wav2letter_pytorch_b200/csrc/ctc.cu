// CTC loss + gradient (log-space alpha/beta over the blank-extended label lattice).
// Replaces nn.CTCLoss(blank=0, reduction='mean', zero_infinity=True) and its backward
// (base_asr_models.py:23, 81, 90).
//
// Pipeline (one stream, no host sync):
//   ctc_prep_kernel   : log_softmax (when given logits) -> lp2 = log2-domain log-probs, rows padded to Cp
//   ctc_alpha_kernel  : one CTA per utterance; each WARP owns a block of 32*R lattice states held in
//                       registers (R per lane); neighbours along the label axis travel by warp shuffle,
//                       warp-boundary states through shared memory; frames of lp2 are staged in shared
//                       memory by a cp.async ring; alpha rows are spilled once to the workspace
//   ctc_beta_grad_kernel : the mirrored beta recursion fused with the gradient: per frame, occupancies
//                       gamma = 2^(alpha+beta-lp-ll) are binned per class in shared memory and the row
//                       softmax - occupancy is emitted (identical for log-prob and logit inputs, because
//                       log_softmax backward is the identity on it -- SURVEY 8a-7)
//   ctc_finish_kernel : per-utterance nll and the reduced loss (deterministic order)
// Every R_RENORM frames the lattice is re-centred on its maximum and the offsets are carried in fp64, so
// fp32 rounding does not grow with the utterance length.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace w2l {

constexpr float kNeg = -1.0e30f;            // finite stand-in for log(0)
constexpr float kDead = -1.0e29f;           // anything below is "log(0)"
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kRing = 8;                     // frames of look-ahead in the cp.async rings
constexpr int kRenorm = 32;
// Occupancies (each in [0,1], summing to <= 1 per class and frame) are accumulated as unsigned fixed point:
// shared-memory integer atomics are native (fp32 ones compile to CAS loops) and make the gradient deterministic.
constexpr float kFix = 1073741824.f;         // 2^30

// Raw MUFU ops (no denormal fix-up code on the serial critical path): arguments are <= 0, results in (0, 1] resp. [0, log2 3].
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lse2_2(float a, float b) {
  const float m = fmaxf(a, b);
  return m + fast_lg2(fast_ex2(a - m) + fast_ex2(b - m));
}
__device__ __forceinline__ float lse2_3(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  return m + fast_lg2(fast_ex2(a - m) + fast_ex2(b - m) + fast_ex2(c - m));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ------------------------------------------------------------------------------------------------
// prep: one warp per (n, t) row.
__global__ void ctc_prep_kernel(const float* __restrict__ x, int from_logits, int N, int T, int C, int64_t stride_n,
                                int64_t stride_t, const int32_t* __restrict__ in_len, float* __restrict__ lp2, int Cp) {
  const int warps_per_block = blockDim.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  if (row >= (int64_t)N * T) return;
  const int n = (int)(row / T), t = (int)(row - (int64_t)n * T);
  const int Tn = max(0, min(T, in_len[n]));
  if (t >= Tn) return;
  const int lane = threadIdx.x & 31;
  const float* src = x + (int64_t)n * stride_n + (int64_t)t * stride_t;
  float* dst = lp2 + row * Cp;
  float lse = 0.f;
  if (from_logits) {
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, src[c]);
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += __expf(src[c] - m);
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    lse = m + __logf(s);
  }
  for (int c = lane; c < Cp; c += 32) {
    float v = kNeg;
    if (c < C) v = fmaxf((src[c] - lse) * kLog2e, kNeg);
    dst[c] = v;
  }
}

// ------------------------------------------------------------------------------------------------
struct CtcMeta {          // per-utterance results of the alpha pass (workspace)
  double ll2;             // log2 likelihood (valid when feasible)
  int32_t feasible;
  int32_t pad;
};

template <int R>
struct LaneLabels {
  int lab[R / 2];          // label of local state 2j+1
  unsigned skip;           // bit j: the s-2 -> s transition is allowed into local state 2j+1
  unsigned skip_next;      // bit j: the s -> s+2 transition is allowed out of local state 2j+1 (beta)
};

template <int R>
__device__ __forceinline__ void load_labels(LaneLabels<R>& lb, const int32_t* __restrict__ tg, int S, int s0, int dead_col) {
  lb.skip = 0;
  lb.skip_next = 0;
#pragma unroll
  for (int j = 0; j < R / 2; ++j) {
    int i = (s0 >> 1) + j;  // label index of state s0 + 2j + 1
    int l = dead_col;       // states past the lattice read a column that holds log(0): they stay dead without a select
    if (i < S) {
      l = tg[i];
      if (i > 0 && l != tg[i - 1]) lb.skip |= 1u << j;
      if (i + 1 < S && tg[i + 1] != l) lb.skip_next |= 1u << j;
    }
    lb.lab[j] = l;
  }
}

// Block-wide maximum of `v` (all threads get it).  s_red has >= 32 floats.
__device__ __forceinline__ float block_max(float v, float* s_red) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float m = kNeg;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, s_red[w]);
  return m;
}

// ------------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(1024)
ctc_alpha_kernel(const float* __restrict__ lp2, int T, int Cp, const int32_t* __restrict__ targets, int64_t tstride,
                 const int32_t* __restrict__ in_len, const int32_t* __restrict__ tg_len, int blank,
                 float* __restrict__ alpha_ws, double* __restrict__ alpha_off, CtcMeta* __restrict__ meta, int Lp, int dbg) {
  extern __shared__ __align__(16) float smem[];
  float* ring = smem;                                   // [kRing][Cp]
  float* s_bnd = ring + kRing * Cp;                     // [2][warps][2]
  float* s_red = s_bnd + 2 * 32 * 2;                    // [32]
  float* s_fin = s_red + 32;                            // [2]
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Tn = max(0, min(T, in_len[n]));
  const int S = max(0, min((int)tstride, tg_len[n]));
  const int L = 2 * S + 1;
  const int s0 = tid * R;
  if (Tn == 0) {
    if (tid == 0) {
      meta[n].ll2 = 0.0;
      meta[n].feasible = (S == 0);
    }
    return;
  }
  LaneLabels<R> lb;
  load_labels<R>(lb, targets + (int64_t)n * tstride, S, s0, Cp - 1);
  const float* lp_n = lp2 + (int64_t)n * T * Cp;
  float* arow = alpha_ws + (int64_t)n * T * Lp + s0;
  double* aoff = alpha_off + (int64_t)n * T;

  // prologue: prefetch rows 0..kRing-2
  const int cp_lanes = Cp >> 2;
#pragma unroll
  for (int p = 0; p < kRing - 1; ++p) {
    if (p < Tn && tid < cp_lanes) cp_async16(ring + p * Cp + tid * 4, lp_n + (int64_t)p * Cp + tid * 4);
    cp_async_commit();
  }
  cp_async_wait<kRing - 2>();
  __syncthreads();

  float a[R];
#pragma unroll
  for (int r = 0; r < R; ++r) a[r] = kNeg;
  if (tid == 0) {
    a[0] = ring[blank];
    if (L > 1) a[1] = ring[lb.lab[0]];
  }
  double off = 0.0;
#pragma unroll
  for (int r = 0; r < R; ++r) arow[r] = a[r];
  if (tid == 0) aoff[0] = 0.0;
  if (lane == 31) {
    s_bnd[(0 * 32 + warp) * 2 + 0] = a[R - 1];
    s_bnd[(0 * 32 + warp) * 2 + 1] = a[R - 2];
  }
  {  // keep the ring full
    int p = kRing - 1;
    if (p < Tn && tid < cp_lanes) cp_async16(ring + (p % kRing) * Cp + tid * 4, lp_n + (int64_t)p * Cp + tid * 4);
    cp_async_commit();
    cp_async_wait<kRing - 2>();
  }
  __syncthreads();

  for (int t = 1; t < Tn; ++t) {
    const float* row = ring + (t % kRing) * Cp;
    // R is even, so a lane's first state is a blank (needs s-1 only) and its second a label whose s-2
    // is the previous lane's LAST state: one value (p1) crosses the lane boundary per frame.
    float p1 = __shfl_up_sync(0xffffffffu, a[R - 1], 1);
    if (lane == 0) p1 = (warp == 0) ? kNeg : s_bnd[(((t - 1) & 1) * 32 + warp - 1) * 2 + 0];
    const float lpb = row[blank];
#pragma unroll
    for (int r = R - 1; r >= 0; --r) {
      const float am1 = (r >= 1) ? a[(r >= 1) ? r - 1 : 0] : p1;
      float v;
      if (r & 1) {
        const float am2 = (r >= 2) ? a[(r >= 2) ? r - 2 : 0] : p1;
        const float sk = ((lb.skip >> (r >> 1)) & 1u) ? am2 : kNeg;
        v = lse2_3(a[r], am1, sk) + row[lb.lab[r >> 1]];
      } else {
        v = lse2_2(a[r], am1) + lpb;
      }
      a[r] = v;
    }
    if ((t % kRenorm) == 0) {   // re-centre on the lattice maximum (uniform branch)
      float m = a[0];
#pragma unroll
      for (int r = 1; r < R; ++r) m = fmaxf(m, a[r]);
      m = block_max(m, s_red);
      if (m > kDead) {
#pragma unroll
        for (int r = 0; r < R; ++r) a[r] = fmaxf(a[r] - m, kNeg);
        off += (double)m;
      }
    }
    if (lane == 31) {
      s_bnd[((t & 1) * 32 + warp) * 2 + 0] = a[R - 1];
      s_bnd[((t & 1) * 32 + warp) * 2 + 1] = a[R - 2];
    }
    if (!(dbg & 1)) {
      float* dst = arow + (int64_t)t * Lp;
#pragma unroll
      for (int r = 0; r < R; ++r) dst[r] = a[r];
      if (tid == 0) aoff[t] = off;
    }
    if (!(dbg & 2)) {
      int p = t + kRing - 1;
      if (p < Tn && tid < cp_lanes) cp_async16(ring + (p % kRing) * Cp + tid * 4, lp_n + (int64_t)p * Cp + tid * 4);
      cp_async_commit();
      cp_async_wait<kRing - 2>();
    }
    if (!(dbg & 4)) __syncthreads();
  }
  // log-likelihood: lse(alpha[L-1], alpha[L-2])
  if (tid < 2) s_fin[tid] = kNeg;
  __syncthreads();
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (s0 + r == L - 1) s_fin[0] = a[r];
    if (s0 + r == L - 2) s_fin[1] = a[r];
  }
  __syncthreads();
  if (tid == 0) {
    float ll = lse2_2(s_fin[0], s_fin[1]);
    meta[n].feasible = ll > kDead;
    meta[n].ll2 = off + (double)ll;
  }
}

// ------------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(1024)
ctc_beta_grad_kernel(const float* __restrict__ lp2, int T, int C, int Cp, const int32_t* __restrict__ targets,
                     int64_t tstride, const int32_t* __restrict__ in_len, const int32_t* __restrict__ tg_len, int blank,
                     const float* __restrict__ alpha_ws, const double* __restrict__ alpha_off,
                     const CtcMeta* __restrict__ meta, int Lp, int zero_infinity, int reduction_mean, int N,
                     float* __restrict__ grad) {
  extern __shared__ __align__(16) float smem[];
  const int nthreads = blockDim.x;
  float* ring = smem;                                   // [kRing][Cp]
  float* aring = ring + kRing * Cp;                     // [kRing][nthreads*R]
  float* s_bnd = aring + kRing * nthreads * R;          // [2][32][2]
  float* s_red = s_bnd + 2 * 32 * 2;                    // [32]
  uint32_t* bins = reinterpret_cast<uint32_t*>(s_red + 32);   // [2][Cp] occupancy sums, fixed point 2^-30
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Tn = max(0, min(T, in_len[n]));
  const int S = max(0, min((int)tstride, tg_len[n]));
  const int L = 2 * S + 1;
  const int s0 = tid * R;
  float* g_n = grad + (int64_t)n * T * C;
  const bool feasible = meta[n].feasible != 0;
  // rows past the utterance (and everything for an infeasible utterance) are exact zeros
  {
    const int first = (feasible && Tn > 0) ? Tn : 0;
    const float fill = (!feasible && !zero_infinity) ? NAN : 0.f;
    for (int64_t i = (int64_t)first * C + tid; i < (int64_t)T * C; i += nthreads) g_n[i] = (i < (int64_t)Tn * C) ? fill : 0.f;
  }
  if (!feasible || Tn == 0) return;
  const float gscale = reduction_mean ? 1.f / ((float)N * (float)max(S, 1)) : 1.f;
  const double ll2 = meta[n].ll2;

  LaneLabels<R> lb;
  load_labels<R>(lb, targets + (int64_t)n * tstride, S, s0, Cp - 1);
  const float* lp_n = lp2 + (int64_t)n * T * Cp;
  const float* arow = alpha_ws + (int64_t)n * T * Lp + s0;
  const double* aoff = alpha_off + (int64_t)n * T;
  const int cp_lanes = Cp >> 2;

  auto prefetch = [&](int p) {   // frame p -> ring slot p % kRing (lp2 row + this thread's alpha states)
    if (p >= 0) {
      if (tid < cp_lanes) cp_async16(ring + (p % kRing) * Cp + tid * 4, lp_n + (int64_t)p * Cp + tid * 4);
      float* adst = aring + (p % kRing) * nthreads * R + s0;
      const float* asrc = arow + (int64_t)p * Lp;
      if (R % 4 == 0) {
#pragma unroll
        for (int r = 0; r < R; r += 4) cp_async16(adst + r, asrc + r);
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) cp_async4(adst + r, asrc + r);
      }
    }
    cp_async_commit();
  };
  for (int i = tid; i < 2 * Cp; i += nthreads) bins[i] = 0u;
#pragma unroll
  for (int q = 0; q < kRing - 1; ++q) prefetch(Tn - 1 - q);
  cp_async_wait<kRing - 2>();
  __syncthreads();

  float b[R];
  double off = 0.0;
  float soft_prev = 0.f;   // softmax value of the row emitted one step later (threads tid < C)

  double aoff_t = aoff[Tn - 1], aoff_next = 0.0;   // software-pipelined: the load for t-1 is issued at step t
  for (int t = Tn - 1; t >= 0; --t) {
    const float* row = ring + (t % kRing) * Cp;
    const float lpb = row[blank];
    if (t > 0) aoff_next = aoff[t - 1];
    if (t == Tn - 1) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int s = s0 + r;
        float v = kNeg;
        if (s == L - 1) v = lpb;
        if (s == L - 2 && (r & 1)) v = row[lb.lab[r >> 1]];
        b[r] = v;
      }
    } else {
      float n1 = __shfl_down_sync(0xffffffffu, b[0], 1);
      float n2 = __shfl_down_sync(0xffffffffu, b[1], 1);
      if (lane == 31) {
        if (warp == (nthreads >> 5) - 1) {
          n1 = kNeg;
          n2 = kNeg;
        } else {
          n1 = s_bnd[(((t + 1) & 1) * 32 + warp + 1) * 2 + 0];
          n2 = s_bnd[(((t + 1) & 1) * 32 + warp + 1) * 2 + 1];
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float bp1 = (r + 1 < R) ? b[(r + 1 < R) ? r + 1 : 0] : n1;
        float v;
        if (r & 1) {
          // s odd: s+2 is the next label state; r+2 < R -> own register, else next lane (n2 when r == R-1)
          const float bp2 = (r + 2 < R) ? b[(r + 2 < R) ? r + 2 : 0] : n2;
          const float sk = ((lb.skip_next >> (r >> 1)) & 1u) ? bp2 : kNeg;
          v = lse2_3(b[r], bp1, sk) + row[lb.lab[r >> 1]];
        } else {
          v = lse2_2(b[r], bp1) + lpb;
        }
        b[r] = v;
      }
      if (((Tn - 1 - t) % kRenorm) == 0) {
        float m = b[0];
#pragma unroll
        for (int r = 1; r < R; ++r) m = fmaxf(m, b[r]);
        m = block_max(m, s_red);
        if (m > kDead) {
#pragma unroll
          for (int r = 0; r < R; ++r) b[r] = fmaxf(b[r] - m, kNeg);
          off += (double)m;
        }
      }
    }
    if (lane == 0) {
      s_bnd[((t & 1) * 32 + warp) * 2 + 0] = b[0];
      s_bnd[((t & 1) * 32 + warp) * 2 + 1] = b[1];
    }
    // occupancies gamma_t(s) = 2^(alpha + beta - lp - ll) binned per class
    {
      const float cst = (float)(aoff_t + off - ll2);
      aoff_t = aoff_next;
      const float* al = aring + (t % kRing) * nthreads * R + s0;
      uint32_t* bin = bins + (t & 1) * Cp;
      float gb = 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (r & 1) {
          const int l = lb.lab[r >> 1];
          const float g = exp2f(al[r] + b[r] - row[l] + cst);
          if (s0 + r < L && g > 0.f) atomicAdd(bin + l, __float2uint_rn(fminf(g, 1.5f) * kFix));
        } else {
          const float g = exp2f(al[r] + b[r] - lpb + cst);
          gb += (s0 + r < L) ? g : 0.f;
        }
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) gb += __shfl_xor_sync(0xffffffffu, gb, o);
      if (lane == 0 && gb > 0.f) atomicAdd(bin + blank, __float2uint_rn(fminf(gb, 1.5f) * kFix));
    }
    // emit the row completed at the previous step (t+1), then remember this row's softmax
    if (tid < C) {
      if (t < Tn - 1) {
        uint32_t* bprev = bins + ((t + 1) & 1) * Cp;
        g_n[(int64_t)(t + 1) * C + tid] = (soft_prev - (float)bprev[tid] * (1.f / kFix)) * gscale;
        bprev[tid] = 0u;
      }
      soft_prev = exp2f(row[tid]);
    }
    prefetch(t - (kRing - 1));
    cp_async_wait<kRing - 2>();
    __syncthreads();
  }
  if (tid < C) g_n[tid] = (soft_prev - (float)bins[tid] * (1.f / kFix)) * gscale;   // row 0 (parity 0)
}

__global__ void ctc_finish_kernel(const CtcMeta* __restrict__ meta, const int32_t* __restrict__ tg_len, int64_t tstride, int N,
                                  int zero_infinity, int reduction_mean, float* __restrict__ nll, float* __restrict__ loss) {
  __shared__ double s_sum[256];
  double acc = 0.0;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float v;
    if (meta[n].feasible) v = (float)(-meta[n].ll2 * (double)kLn2);
    else v = zero_infinity ? 0.f : INFINITY;
    if (nll) nll[n] = v;
    const int S = max(0, min((int)tstride, tg_len[n]));
    acc += reduction_mean ? (double)v / (double)max(S, 1) : (double)v;
  }
  s_sum[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0 && loss) {
    double tot = 0.0;
    for (int i = 0; i < (int)blockDim.x; ++i) tot += s_sum[i];   // fixed order => deterministic
    loss[0] = (float)(reduction_mean ? tot / (double)N : tot);
  }
}

// ---------------------------------------------------------------- host side
struct CtcPlan {
  int R, threads, Lp, Cp;
  size_t off_lp2, off_alpha, off_aoff, off_meta, total;
};

static bool make_plan(int64_t N, int64_t T, int64_t S_max, int64_t C, CtcPlan* p) {
  const int64_t L = 2 * S_max + 1;
  int R = 2;
  while (R <= 8 && (L + 32 * R - 1) / (32 * R) > 16) R *= 2;
  if (R > 8) {
    R = 8;
    if ((L + 32 * R - 1) / (32 * R) > 32) return false;
  }
  p->R = R;
  p->threads = (int)((L + 32 * R - 1) / (32 * R)) * 32;
  p->Lp = p->threads * R;
  p->Cp = (int)((C + 32) / 32) * 32;            // > C: the last column always holds log(0) (dead lattice states read it)
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t at = o;
    o += (bytes + 255) & ~(size_t)255;
    return at;
  };
  p->off_lp2 = take((size_t)N * T * p->Cp * sizeof(float));
  p->off_alpha = take((size_t)N * T * p->Lp * sizeof(float));
  p->off_aoff = take((size_t)N * T * sizeof(double));
  p->off_meta = take((size_t)N * sizeof(CtcMeta));
  p->total = o;
  return true;
}

template <int R>
static int launch_ctc(const CtcPlan& pl, char* ws, int64_t N, int64_t T, int64_t C, const int32_t* targets, int64_t tstride,
                      const int32_t* in_len, const int32_t* tg_len, int blank, int zero_infinity, int reduction_mean, float* grad,
                      cudaStream_t st) {
  float* lp2 = (float*)(ws + pl.off_lp2);
  float* alpha = (float*)(ws + pl.off_alpha);
  double* aoff = (double*)(ws + pl.off_aoff);
  CtcMeta* meta = (CtcMeta*)(ws + pl.off_meta);
  const size_t smem_a = (size_t)(kRing * pl.Cp + 2 * 32 * 2 + 32 + 2) * sizeof(float);
  const size_t smem_b = (size_t)(kRing * pl.Cp + kRing * pl.threads * R + 2 * 32 * 2 + 32 + 2 * pl.Cp) * sizeof(float);
  W2L_REQUIRE(smem_b <= 220 * 1024, "ctc: shared memory %zu too large", smem_b);
  W2L_CUDA(cudaFuncSetAttribute(ctc_alpha_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
  W2L_CUDA(cudaFuncSetAttribute(ctc_beta_grad_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
  static const int dbg = getenv("W2L_CTC_DBG") ? atoi(getenv("W2L_CTC_DBG")) : 0;      // development knob (timing experiments)
  ctc_alpha_kernel<R><<<(unsigned)N, pl.threads, smem_a, st>>>(lp2, (int)T, pl.Cp, targets, tstride, in_len, tg_len, blank, alpha,
                                                             aoff, meta, pl.Lp, dbg);
  int rc = after_launch("ctc_alpha_kernel");
  if (rc) return rc;
  if (grad) {
    ctc_beta_grad_kernel<R><<<(unsigned)N, pl.threads, smem_b, st>>>(lp2, (int)T, (int)C, pl.Cp, targets, tstride, in_len, tg_len,
                                                                   blank, alpha, aoff, meta, pl.Lp, zero_infinity, reduction_mean,
                                                                   (int)N, grad);
    rc = after_launch("ctc_beta_grad_kernel");
  }
  return rc;
}

}  // namespace w2l

extern "C" {

size_t w2l_ctc_loss_workspace_bytes(int64_t N, int64_t T, int64_t S_max) {
  w2l::CtcPlan p;
  // class count only affects the lp2 staging; size it for the largest supported alphabet row (128)
  if (N <= 0 || T <= 0) return 256;
  if (!w2l::make_plan(N, T, S_max < 0 ? 0 : S_max, 128, &p)) return 0;
  return p.total;
}

int w2l_ctc_loss(const float* x, int32_t from_logits, int64_t N, int64_t T, int64_t C, int64_t stride_n, int64_t stride_t,
                 const int32_t* targets, int64_t target_stride, const int32_t* input_lengths, const int32_t* target_lengths,
                 int32_t blank, int32_t zero_infinity, int32_t reduction_mean, float* nll, float* grad, float* loss,
                 void* workspace, size_t workspace_bytes, void* stream) {
  using namespace w2l;
  W2L_REQUIRE(N >= 1 && T >= 1 && C >= 2 && C <= 128, "ctc_loss: unsupported shape N=%lld T=%lld C=%lld (need C<=128)", (long long)N,
              (long long)T, (long long)C);
  W2L_REQUIRE(x && input_lengths && target_lengths && workspace, "ctc_loss: null pointer");
  W2L_REQUIRE(target_stride == 0 || targets, "ctc_loss: null targets");
  W2L_REQUIRE(blank >= 0 && blank < C, "ctc_loss: blank out of range");
  CtcPlan pl;
  W2L_REQUIRE(make_plan(N, T, target_stride, C, &pl), "ctc_loss: target length %lld exceeds the supported lattice (4095 labels)",
              (long long)target_stride);
  W2L_REQUIRE(workspace_bytes >= pl.total, "ctc_loss: workspace too small (%zu < %zu)", workspace_bytes, pl.total);
  W2L_REQUIRE(((uintptr_t)workspace & 255) == 0, "ctc_loss: workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  {
    const int64_t rows = N * T;
    const int wpb = 8;
    ctc_prep_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, st>>>(x, from_logits, (int)N, (int)T, (int)C, stride_n,
                                                                           stride_t, input_lengths, (float*)(ws + pl.off_lp2),
                                                                           pl.Cp);
    int rc = after_launch("ctc_prep_kernel");
    if (rc) return rc;
  }
  int rc;
  switch (pl.R) {
    case 2: rc = launch_ctc<2>(pl, ws, N, T, C, targets, target_stride, input_lengths, target_lengths, blank, zero_infinity, reduction_mean, grad, st); break;
    case 4: rc = launch_ctc<4>(pl, ws, N, T, C, targets, target_stride, input_lengths, target_lengths, blank, zero_infinity, reduction_mean, grad, st); break;
    default: rc = launch_ctc<8>(pl, ws, N, T, C, targets, target_stride, input_lengths, target_lengths, blank, zero_infinity, reduction_mean, grad, st); break;
  }
  if (rc) return rc;
  if (nll || loss) {
    ctc_finish_kernel<<<1, 256, 0, st>>>((const CtcMeta*)(ws + pl.off_meta), target_lengths, target_stride, (int)N, zero_infinity,
                                         reduction_mean, nll, loss);
    rc = after_launch("ctc_finish_kernel");
  }
  return rc;
}

}  // extern "C"
