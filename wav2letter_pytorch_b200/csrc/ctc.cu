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
// log2-sum-exp2 with the maximum's term folded to the constant 1: one MUFU op fewer than exponentiating every argument
// (the recursion is MUFU- and latency-bound: 2 / 3 instead of 3 / 4 special-function ops per lattice state)
__device__ __forceinline__ float lse2_2(float a, float b) {
  const float m = fmaxf(a, b);
  return m + fast_lg2(1.f + fast_ex2(fminf(a, b) - m));
}
__device__ __forceinline__ float lse2_3(float a, float b, float c) {
  const float hi = fmaxf(a, b), lo = fminf(a, b);
  const float m = fmaxf(hi, c);
  const float mid = fmaxf(lo, fminf(hi, c)), low = fminf(lo, c);
  return m + fast_lg2(1.f + fast_ex2(mid - m) + fast_ex2(low - m));
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
__device__ __forceinline__ void
ctc_alpha_body(float* smem, int n, const float* __restrict__ lp2, int T, int Cp, const int32_t* __restrict__ targets, int64_t tstride,
               const int32_t* __restrict__ in_len, const int32_t* __restrict__ tg_len, int blank,
               float* __restrict__ alpha_ws, double* __restrict__ alpha_off, CtcMeta* __restrict__ meta, int Lp, int dbg) {
  float* ring = smem;                                   // [kRing][Cp]
  float* s_bnd = ring + kRing * Cp;                     // [2][warps][2]
  float* s_red = s_bnd + 2 * 32 * 2;                    // [32]
  float* s_fin = s_red + 32;                            // [2]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x;
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
    if (p < Tn)
      for (int q = tid; q < cp_lanes; q += nthreads) cp_async16(ring + p * Cp + q * 4, lp_n + (int64_t)p * Cp + q * 4);
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
    if (p < Tn)
      for (int q = tid; q < cp_lanes; q += nthreads) cp_async16(ring + (p % kRing) * Cp + q * 4, lp_n + (int64_t)p * Cp + q * 4);
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
      if (p < Tn)
        for (int q = tid; q < cp_lanes; q += nthreads) cp_async16(ring + (p % kRing) * Cp + q * 4, lp_n + (int64_t)p * Cp + q * 4);
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


// ---- parallel schedule: alpha and beta recursions in separate CTAs, both spilled, gradient in a third, fully parallel pass.
// The recursion is a serial chain of T dependent steps, so everything in the per-frame body is trimmed to the chain:
//   * frames are processed in blocks of kBlk: the block's lp rows arrive by cp.async (double-buffered), the lattice is re-centred
//     on its maximum once per block (offsets carried in fp64, one per block), the CTA synchronises twice per block;
//   * inside a block there is NO CTA barrier: the one (alpha) / two (beta) lattice values that cross a warp boundary per frame
//     travel through a shared-memory slot tagged with the frame number, polled by the consuming lane -- lower warps run ahead,
//     so the poll normally hits at once (a wavefront over the warps);
//   * skip-transition masks are additive (0 / -inf) constants, spills are one vector store per thread.
constexpr int kBlk = 32;

template <int R>
__device__ __forceinline__ void spill_states(float* dst, const float (&v)[R]) {
  if (R == 2) {
    *reinterpret_cast<float2*>(dst) = make_float2(v[0], v[1]);
  } else {
#pragma unroll
    for (int r = 0; r < R; r += 4) *reinterpret_cast<float4*>(dst + r) = make_float4(v[r], v[r + 1], v[r + 2], v[r + 3]);
  }
}

__device__ __forceinline__ void slot_put(float4* slot, float v0, float v1, int tag, float w = 0.f) {
  asm volatile("st.volatile.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(slot)), "f"(v0), "f"(v1), "f"(__int_as_float(tag)), "f"(w)
               : "memory");
}
// Every lane of the warp reads the same address (a broadcast), so the retry branch is warp-uniform.  Tags only grow, and the
// slot for step i holds tag i or (stale) i - kBlk: "tag >= wanted" means ready.  The load for step i is issued one iteration
// early (slot_load), when the feeding warp -- one step ahead in the wavefront -- has normally published it already, so its
// latency hides behind the current step's arithmetic; slot_ready re-polls only if that speculative read came too soon.
__device__ __forceinline__ float4 slot_load(const float4* slot) {
  float4 q;
  asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(smem_u32(slot)) : "memory");
  return q;
}
__device__ __forceinline__ void slot_ready(float4& q, const float4* slot, int tag) {
  uint32_t spins = 0;
  while (__float_as_int(q.z) < tag) {
    q = slot_load(slot);
    if (++spins > (1u << 24)) __trap();                 // protocol bug: trap instead of hanging the GPU
  }
}

// One frame of the recursion on a thread's R states.  x1 (, x2): the neighbouring lane's boundary value(s) of the previous frame.
template <int R, bool BETA>
__device__ __forceinline__ void lattice_step(float (&v)[R], float x1, float x2, float lpb, const float (&lpl)[R / 2],
                                             const float (&skadd)[R / 2]) {
  if (!BETA) {
    // a lane's first state is a blank (needs s-1), its second a label whose s-2 is the previous lane's LAST state
#pragma unroll
    for (int r = R - 1; r >= 0; --r) {
      const float am1 = (r >= 1) ? v[(r >= 1) ? r - 1 : 0] : x1;
      if (r & 1) {
        const float am2 = (r >= 2) ? v[(r >= 2) ? r - 2 : 0] : x1;
        v[r] = lse2_3(v[r], am1, am2 + skadd[r >> 1]) + lpl[r >> 1];
      } else {
        v[r] = lse2_2(v[r], am1) + lpb;
      }
    }
  } else {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float bp1 = (r + 1 < R) ? v[(r + 1 < R) ? r + 1 : 0] : x1;
      if (r & 1) {
        const float bp2 = (r + 2 < R) ? v[(r + 2 < R) ? r + 2 : 0] : x2;
        v[r] = lse2_3(v[r], bp1, bp2 + skadd[r >> 1]) + lpl[r >> 1];
      } else {
        v[r] = lse2_2(v[r], bp1) + lpb;
      }
    }
  }
}

// One direction of the lattice recursion for utterance n.  BETA = false: alpha_t(s) = lse(a(s), a(s-1), [skip] a(s-2)) + lp_t(l_s),
// t ascending.  BETA = true: beta_t(s) = lse(b(s), b(s+1), [skip] b(s+2)) + lp_t(l_s), t descending (ATen convention: beta
// includes the emission at t).  Step i is frame t = BETA ? Tn-1-i : i; rows are spilled at ws[n][t][s]; off_blk[n][i / kBlk]
// is the fp64 offset of the values spilled during block i / kBlk.
template <int R, bool BETA>
__device__ __forceinline__ void
lattice_pass(float* smem, int n, const float* __restrict__ lp2, int T, int Cp, const int32_t* __restrict__ targets, int64_t tstride,
             const int32_t* __restrict__ in_len, const int32_t* __restrict__ tg_len, int blank, float* __restrict__ ws,
             double* __restrict__ off_blk, int n_blk, CtcMeta* __restrict__ meta, int Lp) {
  float* lpbuf = smem;                                                    // [2][kBlk][Cp]
  float4* slots = reinterpret_cast<float4*>(lpbuf + 2 * kBlk * Cp);       // [33][kBlk]: one ring per warp + an always-ready dummy ring
  float* s_red = reinterpret_cast<float*>(slots + 33 * kBlk);             // [32]
  float* s_fin = s_red + 32;                                              // [2]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x, nwarps = nthreads >> 5;
  const int Tn = max(0, min(T, in_len[n]));
  const int S = max(0, min((int)tstride, tg_len[n]));
  const int L = 2 * S + 1;
  const int s0 = tid * R;
  if (Tn == 0) {
    if (!BETA && tid == 0) {
      meta[n].ll2 = 0.0;
      meta[n].feasible = (S == 0);
    }
    return;
  }
  // per-thread lattice constants: column (byte offset in an lp row) of each odd (label) state, additive skip mask
  int lab_b[R / 2];
  float skadd[R / 2];
  {
    const int32_t* tg = targets + (int64_t)n * tstride;
#pragma unroll
    for (int j = 0; j < R / 2; ++j) {
      const int i = (s0 >> 1) + j;                // label index of state s0 + 2j + 1
      int l = Cp - 1;                              // states past the lattice read the column that holds log(0): they stay dead
      bool skip = false;
      if (i < S) {
        l = tg[i];
        skip = BETA ? (i + 1 < S && tg[i + 1] != l) : (i > 0 && tg[i - 1] != l);
      }
      lab_b[j] = l * 4;
      skadd[j] = skip ? 0.f : kNeg;
    }
  }
  const int blank_b = blank * 4;
  for (int i = tid; i < 33 * kBlk; i += nthreads)   // ring 32 feeds the warp that has no neighbour: log(0), tag = "always ready"
    slots[i] = i < 32 * kBlk ? make_float4(0.f, 0.f, __int_as_float(-1), 0.f) : make_float4(kNeg, kNeg, __int_as_float(0x7fffffff), 0.f);
  const float* lp_n = lp2 + (int64_t)n * T * Cp;
  const int blocks = (Tn + kBlk - 1) / kBlk;
  // rows of block k: steps [k*kBlk, k*kBlk + cnt) = frames [f_lo, f_lo + cnt), ascending in memory for both directions
  auto prefetch_block = [&](int k) {
    if (k < blocks) {
      const int i0 = k * kBlk, cnt = min(kBlk, Tn - i0);
      const int f_lo = BETA ? Tn - i0 - cnt : i0;
      const float* src = lp_n + (int64_t)f_lo * Cp;
      float* dst = lpbuf + (k & 1) * kBlk * Cp;
      for (int q = tid; q < cnt * (Cp >> 2); q += nthreads) cp_async16(dst + q * 4, src + q * 4);
    }
    cp_async_commit();
  };
  prefetch_block(0);
  prefetch_block(1);

  float v[R];
#pragma unroll
  for (int r = 0; r < R; ++r) v[r] = kNeg;
  double off = 0.0;
  double* offs = off_blk + (int64_t)n * n_blk;
  const int64_t wstep = BETA ? -(int64_t)Lp : (int64_t)Lp;
  float* wp = ws + ((int64_t)n * T + (BETA ? Tn - 1 : 0)) * Lp + s0;     // spill row of the current step
  const bool has_nb = BETA ? (warp < nwarps - 1) : (warp > 0);            // a neighbouring warp feeds this warp's boundary lane
  const bool edge = BETA ? (lane == 31) : (lane == 0);                    // the lane that takes the neighbouring warp's values
  const bool pub = BETA ? (lane == 0) : (lane == 31);                     // the lane that publishes this warp's boundary values
  const float4* slot_in = slots + (has_nb ? (BETA ? warp + 1 : warp - 1) : 32) * kBlk;
  float4* slot_out = slots + warp * kBlk;
  const int row_b = BETA ? -Cp * 4 : Cp * 4;

  for (int k = 0; k < blocks; ++k) {
    cp_async_wait<1>();                       // block k's rows have landed (block k+1 may still be in flight)
    __syncthreads();
    const int i0 = k * kBlk, cnt = min(kBlk, Tn - i0);
    float shift = 0.f;                        // this block's re-centring, also owed by the boundary values published before it
    if (k > 0) {                              // re-centre the carried lattice on its maximum
      float m = v[0];
#pragma unroll
      for (int r = 1; r < R; ++r) m = fmaxf(m, v[r]);
      m = block_max(m, s_red);
      if (m > kDead) {
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = fmaxf(v[r] - m, kNeg);
        off += (double)m;
        shift = m;
      }
    }
    if (tid == 0) offs[k] = off;
    // first row of the block in step order: ascending frames sit at rows 0.., descending ones at rows cnt-1..
    const char* rowp = reinterpret_cast<const char*>(lpbuf + (k & 1) * kBlk * Cp) + (BETA ? (cnt - 1) * Cp * 4 : 0);
    int ii = 0;
    if (k == 0) {                             // step 0: initial lattice column
      const float lpb = *reinterpret_cast<const float*>(rowp + blank_b);
      if (!BETA) {
        if (tid == 0) {
          v[0] = lpb;
          if (L > 1) v[1] = *reinterpret_cast<const float*>(rowp + lab_b[0]);
        }
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int s = s0 + r;
          if (s == L - 1) v[r] = lpb;
          if (s == L - 2 && (r & 1)) v[r] = *reinterpret_cast<const float*>(rowp + lab_b[r >> 1]);
        }
      }
      if (pub) slot_put(slot_out, v[BETA ? 0 : R - 1], BETA ? v[1] : 0.f, 0);
      spill_states<R>(wp, v);
      wp += wstep;
      rowp += row_b;
      ii = 1;
    }
    float4 q = slot_load(slot_in + ((ii + kBlk - 1) & (kBlk - 1)));
    for (; ii < cnt; ++ii) {
      const int i = i0 + ii;
      // the neighbouring warp's boundary values of step i-1 (read one iteration ago), then the speculative read for step i
      slot_ready(q, slot_in + ((ii + kBlk - 1) & (kBlk - 1)), i - 1);
      const float y1 = fmaxf(q.x - shift, kNeg);
      const float y2 = BETA ? fmaxf(q.y - shift, kNeg) : kNeg;
      q = slot_load(slot_in + ii);
      shift = 0.f;
      const float lpb = *reinterpret_cast<const float*>(rowp + blank_b);
      float lpl[R / 2];
#pragma unroll
      for (int j = 0; j < R / 2; ++j) lpl[j] = *reinterpret_cast<const float*>(rowp + lab_b[j]);
      float x1, x2 = kNeg;
      if (!BETA) {
        x1 = __shfl_up_sync(0xffffffffu, v[R - 1], 1);
      } else {
        x1 = __shfl_down_sync(0xffffffffu, v[0], 1);
        x2 = __shfl_down_sync(0xffffffffu, v[1], 1);
      }
      x1 = edge ? y1 : x1;
      if (BETA) x2 = edge ? y2 : x2;
      lattice_step<R, BETA>(v, x1, x2, lpb, lpl, skadd);
      if (pub) slot_put(slot_out + ii, v[BETA ? 0 : R - 1], BETA ? v[1] : 0.f, i);
      spill_states<R>(wp, v);
      wp += wstep;
      rowp += row_b;
    }
    __syncthreads();                          // every warp is done with block k's rows and slots
    prefetch_block(k + 2);
  }
  if (!BETA) {                                // log-likelihood: lse(alpha[L-1], alpha[L-2])
    if (tid < 2) s_fin[tid] = kNeg;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (s0 + r == L - 1) s_fin[0] = v[r];
      if (s0 + r == L - 2) s_fin[1] = v[r];
    }
    __syncthreads();
    if (tid == 0) {
      const float ll = lse2_2(s_fin[0], s_fin[1]);
      meta[n].feasible = ll > kDead;
      meta[n].ll2 = off + (double)ll;
    }
  }
}

// One launch, 2N CTAs: CTA n runs utterance n's alpha recursion, CTA N+n its beta recursion -- the two serial chains of T
// frames run side by side (they are independent until the gradient), halving the critical path of the loss+gradient.
template <int R>
__global__ void __launch_bounds__(1024)
ctc_lattice_kernel(const float* __restrict__ lp2, int N, int T, int Cp, const int32_t* __restrict__ targets, int64_t tstride,
                   const int32_t* __restrict__ in_len, const int32_t* __restrict__ tg_len, int blank, float* __restrict__ alpha_ws,
                   double* __restrict__ alpha_off, float* __restrict__ beta_ws, double* __restrict__ beta_off, int n_blk,
                   CtcMeta* __restrict__ meta, int Lp) {
  extern __shared__ __align__(16) float smem[];
  if ((int)blockIdx.x < N)
    lattice_pass<R, false>(smem, blockIdx.x, lp2, T, Cp, targets, tstride, in_len, tg_len, blank, alpha_ws, alpha_off, n_blk, meta, Lp);
  else
    lattice_pass<R, true>(smem, blockIdx.x - N, lp2, T, Cp, targets, tstride, in_len, tg_len, blank, beta_ws, beta_off, n_blk, meta, Lp);
}

template <int R>
__global__ void __launch_bounds__(1024)
ctc_alpha_kernel(const float* __restrict__ lp2, int T, int Cp, const int32_t* __restrict__ targets, int64_t tstride,
                 const int32_t* __restrict__ in_len, const int32_t* __restrict__ tg_len, int blank,
                 float* __restrict__ alpha_ws, double* __restrict__ alpha_off, CtcMeta* __restrict__ meta, int Lp, int dbg) {
  extern __shared__ __align__(16) float smem[];
  ctc_alpha_body<R>(smem, blockIdx.x, lp2, T, Cp, targets, tstride, in_len, tg_len, blank, alpha_ws, alpha_off, meta, Lp, dbg);
}

// Gradient from the spilled lattices, fully parallel over (utterance, frame): one warp per frame bins the occupancies
// gamma_t(s) = 2^(alpha + beta - lp - ll) per class (fixed-point shared-memory integer atomics: deterministic) and emits
// softmax - occupancy.  grid (ceil(T / kGradFrames), N), 8 warps, each warp walks kGradFrames/8 frames.
constexpr int kGradFrames = 32;
__device__ __forceinline__ void
ctc_grad_body(float* smem, const float* __restrict__ lp2, int T, int C, int Cp, const int32_t* __restrict__ targets, int64_t tstride,
              const int32_t* __restrict__ in_len, const int32_t* __restrict__ tg_len, int blank,
              const float* __restrict__ alpha_ws, const double* __restrict__ alpha_off, const float* __restrict__ beta_ws,
              const double* __restrict__ beta_off, int n_blk, const CtcMeta* __restrict__ meta, int Lp, int zero_infinity,
              int reduction_mean, int N, float* __restrict__ grad) {
  const int nwarps = blockDim.x >> 5;
  float* rows = smem;                                               // [nwarps][Cp]
  uint32_t* bins = reinterpret_cast<uint32_t*>(rows + nwarps * Cp);  // [nwarps][Cp]
  uint8_t* lab = reinterpret_cast<uint8_t*>(bins + nwarps * Cp);     // [L] class of every lattice state
  const int n = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Tn = max(0, min(T, in_len[n]));
  const int S = max(0, min((int)tstride, tg_len[n]));
  const int L = 2 * S + 1;
  const bool feasible = meta[n].feasible != 0;
  const int t_begin = blockIdx.x * kGradFrames, t_end = min(T, t_begin + kGradFrames);
  float* g_n = grad + (int64_t)n * T * C;
  if (!feasible || Tn == 0 || t_begin >= Tn) {   // exact zeros past the utterance / for an infeasible one (NaN without zero_infinity)
    const float fill = (!feasible && !zero_infinity) ? NAN : 0.f;
    for (int i = t_begin * C + tid; i < t_end * C; i += blockDim.x) g_n[i] = (i < Tn * C) ? fill : 0.f;
    return;
  }
  const int32_t* tg = targets + (int64_t)n * tstride;
  for (int s = tid; s < L; s += blockDim.x) lab[s] = (uint8_t)((s & 1) ? tg[s >> 1] : blank);
  __syncthreads();
  const float gscale = reduction_mean ? 1.f / ((float)N * (float)max(S, 1)) : 1.f;
  const double ll2 = meta[n].ll2;
  float* row = rows + warp * Cp;
  uint32_t* bin = bins + warp * Cp;
  for (int t = t_begin + warp; t < t_end; t += nwarps) {
    float* g_t = g_n + (int64_t)t * C;
    if (t >= Tn) {
      for (int c = lane; c < C; c += 32) g_t[c] = 0.f;
      continue;
    }
    const float* lp_t = lp2 + ((int64_t)n * T + t) * Cp;
    for (int c = lane; c < Cp; c += 32) {
      row[c] = lp_t[c];
      bin[c] = 0u;
    }
    __syncwarp();
    const float cst = (float)(alpha_off[(int64_t)n * n_blk + t / kBlk] + beta_off[(int64_t)n * n_blk + (Tn - 1 - t) / kBlk] - ll2);
    const float* a_t = alpha_ws + ((int64_t)n * T + t) * Lp;
    const float* b_t = beta_ws + ((int64_t)n * T + t) * Lp;
    float gb = 0.f;
    for (int s = lane; s < L; s += 32) {
      const int l = lab[s];
      const float g = exp2f(a_t[s] + b_t[s] - row[l] + cst);
      if (s & 1) {
        if (g > 0.f) atomicAdd(bin + l, __float2uint_rn(fminf(g, 1.5f) * kFix));
      } else {
        gb += g;
      }
    }
    // blank states: lanes hold disjoint partial sums; fold them in a fixed order
#pragma unroll
    for (int o = 16; o; o >>= 1) gb += __shfl_xor_sync(0xffffffffu, gb, o);
    if (lane == 0 && gb > 0.f) atomicAdd(bin + blank, __float2uint_rn(fminf(gb, 1.5f) * kFix));
    __syncwarp();
    for (int c = lane; c < C; c += 32) g_t[c] = (exp2f(row[c]) - (float)bin[c] * (1.f / kFix)) * gscale;
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256)
ctc_grad_kernel(const float* __restrict__ lp2, int T, int C, int Cp, const int32_t* __restrict__ targets, int64_t tstride,
                const int32_t* __restrict__ in_len, const int32_t* __restrict__ tg_len, int blank,
                const float* __restrict__ alpha_ws, const double* __restrict__ alpha_off, const float* __restrict__ beta_ws,
                const double* __restrict__ beta_off, int n_blk, const CtcMeta* __restrict__ meta, int Lp, int zero_infinity,
                int reduction_mean, int N, float* __restrict__ grad) {
  extern __shared__ __align__(16) float smem[];
  ctc_grad_body(smem, lp2, T, C, Cp, targets, tstride, in_len, tg_len, blank, alpha_ws, alpha_off, beta_ws, beta_off, n_blk, meta, Lp,
                zero_infinity, reduction_mean, N, grad);
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
      for (int q = tid; q < cp_lanes; q += nthreads) cp_async16(ring + (p % kRing) * Cp + q * 4, lp_n + (int64_t)p * Cp + q * 4);
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
  // softmax values of the row emitted one step later: thread tid owns columns tid, tid + nthreads, ... (C <= 128 and a CTA has at
  // least 32 threads, so at most kEmitCols of them; with the usual 29 labels only the first is live)
  constexpr int kEmitCols = 4;
  float soft_prev[kEmitCols] = {0.f, 0.f, 0.f, 0.f};

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
#pragma unroll
    for (int q = 0; q < kEmitCols; ++q) {
      const int c = tid + q * nthreads;
      if (c < C) {
        if (t < Tn - 1) {
          uint32_t* bprev = bins + ((t + 1) & 1) * Cp;
          g_n[(int64_t)(t + 1) * C + c] = (soft_prev[q] - (float)bprev[c] * (1.f / kFix)) * gscale;
          bprev[c] = 0u;
        }
        soft_prev[q] = exp2f(row[c]);
      }
    }
    prefetch(t - (kRing - 1));
    cp_async_wait<kRing - 2>();
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < kEmitCols; ++q) {                                              // row 0 (parity 0)
    const int c = tid + q * nthreads;
    if (c < C) g_n[c] = (soft_prev[q] - (float)bins[c] * (1.f / kFix)) * gscale;
  }
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
  bool parallel;   // alpha and beta in parallel CTAs + a parallel gradient pass (beta is spilled too); else alpha, then beta+grad fused
  size_t off_lp2, off_alpha, off_aoff, off_beta, off_boff, off_meta, total;
};

// the parallel schedule doubles the lattice spill; beyond this size the serial (fused beta+gradient) schedule is used
constexpr size_t kCtcParallelSpillLimit = (size_t)3 << 30;

static bool make_plan(int64_t N, int64_t T, int64_t S_max, int64_t C, CtcPlan* p) {
  const int64_t L = 2 * S_max + 1;
  // one warp per SM sub-partition when the lattice allows it (the recursion is a latency chain: fewer, fatter warps keep the
  // per-frame exchange between warps short and give each warp more independent states to interleave)
  int R = 2;
  while (R < 8 && (L + 32 * R - 1) / (32 * R) > 4) R *= 2;
  if ((L + 32 * R - 1) / (32 * R) > 32) return false;
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
  const size_t lattice = (size_t)N * T * p->Lp * sizeof(float);
  p->parallel = 2 * lattice <= kCtcParallelSpillLimit;
  p->off_alpha = take(lattice);
  p->off_aoff = take((size_t)N * T * sizeof(double));          // serial schedule: one per frame; parallel: one per kBlk frames
  p->off_beta = p->parallel ? take(lattice) : 0;
  p->off_boff = p->parallel ? take((size_t)N * T * sizeof(double)) : 0;
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
  const char* dbg_env = getenv("W2L_CTC_DBG");            // development knob: 1/2/4 timing experiments, 8 = force the serial schedule
  const int dbg = dbg_env ? atoi(dbg_env) : 0;
  W2L_REQUIRE(N <= 65535, "ctc: batch %lld exceeds the grid limit", (long long)N);
  if (pl.parallel && grad && !(dbg & 8)) {
    float* beta = (float*)(ws + pl.off_beta);
    double* boff = (double*)(ws + pl.off_boff);
    const int gw = 8;
    const size_t smem_g = (size_t)gw * pl.Cp * 8 + (size_t)(2 * tstride + 1 + 15) + 16 + pl.Cp;   // rows, bins, lab
    W2L_REQUIRE(smem_g <= 48 * 1024, "ctc: gradient pass shared memory %zu too large", smem_g);
    dim3 grid((unsigned)((T + kGradFrames - 1) / kGradFrames), (unsigned)N);
    const int n_blk = (int)((T + kBlk - 1) / kBlk);
    const size_t smem_l = (size_t)(2 * kBlk * pl.Cp) * sizeof(float) + (size_t)33 * kBlk * sizeof(float4) + (32 + 2) * sizeof(float);
    W2L_CUDA(cudaFuncSetAttribute(ctc_lattice_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l));
    ctc_lattice_kernel<R><<<(unsigned)(2 * N), pl.threads, smem_l, st>>>(lp2, (int)N, (int)T, pl.Cp, targets, tstride, in_len, tg_len,
                                                                        blank, alpha, aoff, beta, boff, n_blk, meta, pl.Lp);
    int rc = after_launch("ctc_lattice_kernel");
    if (rc) return rc;
    ctc_grad_kernel<<<grid, gw * 32, smem_g, st>>>(lp2, (int)T, (int)C, pl.Cp, targets, tstride, in_len, tg_len, blank, alpha, aoff, beta,
                                                   boff, n_blk, meta, pl.Lp, zero_infinity, reduction_mean, (int)N, grad);
    return after_launch("ctc_grad_kernel");
  }
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
