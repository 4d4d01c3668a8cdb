// Conv1d forward / backward-data / backward-weight as tcgen05 implicit GEMMs (sm_100a).
// Replaces nn.Conv1d fwd+bwd at wav2letter.py:35-36,42 and jasper.py:96-105,127,433,468.
//
// One persistent, warp-specialised kernel template (192 threads, 1 CTA per SM):
//   warp 0      TMA producer   : cp.async.bulk.tensor (3-D tiled maps, 128B swizzle) into a 4-stage smem ring
//   warp 1      MMA issuer     : one elected thread issues tcgen05.mma (128 x BN x 16, bf16 -> fp32) into one of
//                                two 256-column TMEM accumulators; tcgen05.commit releases smem stages / signals
//                                the epilogue
//   warps 2..5  epilogue       : tcgen05.ld the finished accumulator (overlapping the next tile's mainloop),
//                                fused bias / BatchNorm-fold / ReLU-clamp, bf16 or fp32 stores (or fp32 atomics
//                                for split-K weight gradients)
// The convolution is never unfolded in memory: a K-step is (tap j, 64-channel chunk); tap j just shifts the row
// coordinate of the activation tile by j*dilation, and rows outside the tensor are zero-filled by TMA (which
// is Jasper's zero padding; Wav2Letter's reflection halo is materialised by the producer of the activation).
//   FWD   : D[t, co]  = sum_{j,ci} X[b, t+off+j*d, ci] * W[j, co, ci]      A = X  (K-major)   B = W[j] (K-major)
//   DGRAD : D[u, ci]  = sum_{j,co} dY[b, u-off-j*d, co] * W[j, co, ci]     A = dY (K-major)   B = W[j] (MN-major)
//   WGRAD : D[co, ci] = sum_{b,t}  dY[b, t, co] * X[b, t+off+j*d, ci]      A = dY (MN-major)  B = X    (MN-major)
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace w2l {

enum { MODE_FWD = 0, MODE_DGRAD = 1, MODE_WGRAD = 2 };

constexpr int kStages = 4;
constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                     // bf16 elements per K-step = one 128-byte swizzle row
constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KB
constexpr int kBBytesMax = 256 * kBlockK * 2;   // 32 KB
constexpr int kStageBytes = kABytes + kBBytesMax;
constexpr int kAccCols = 256;
constexpr int kGemmThreads = 192;
constexpr int kAffBytes = 2 * 256 * 8 + 2 * 256 * 4;   // per-accumulator (scale, shift) [+ mean: the fused BatchNorm-backward reduction] of the tile's columns
constexpr size_t kGemmSmem = (size_t)kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ + kAffBytes;
constexpr int kSlabBufs = 3;                    // slab mode: the next chunk's slab is requested a whole chunk ahead
constexpr size_t kGemmSmemMax = 227 * 1024;     // opt-in limit per CTA on sm_100

struct GemmParams {
  CUtensorMap tmA, tmB;
  int32_t B;            // utterances
  int32_t m_tiles;      // tiles along M per utterance (fwd/dgrad) or along Cout (wgrad)
  int32_t n_tiles;
  int32_t BN;
  int32_t k, dil;
  int32_t kc_steps;     // 64-wide chunks of the contraction channels (fwd/dgrad) or of T_out (wgrad)
  int32_t a_row_off;    // fwd/dgrad: row coordinate of tap 0 for output row 0
  int32_t a_tap_step;   // +dil (fwd) / -dil (dgrad)
  int32_t b_row_off;    // wgrad: x row for (t=0, tap 0)
  int32_t M_valid;      // rows to store per utterance (fwd/dgrad) / Cout (wgrad)
  int32_t N_valid;      // columns to store
  int32_t num_tiles;
  // epilogue
  const float* bias;
  const float* scale;
  const float* shift;
  float* bn_stats;          // fwd: [2*N_valid] per-channel sum / sum of squares of the STORED (bf16-rounded) output, or null
  int32_t act;
  int32_t y_dtype;
  void* y;
  int64_t y_batch_stride;   // elements
  int32_t y_row_off;
  int32_t ldy;
  // wgrad
  int32_t splits;           // > 1: some tiles are shared between CTAs (dw pre-zeroed by the caller)
  int32_t wg_rounds;        // wgrad: whole-tile rounds per CTA before the stream-K remainder
  int32_t mn4d;             // wgrad: operands are loaded through 4-D maps (64 x rows x chunks x B), one TMA op each
  int32_t dbg;              // development knobs (env W2L_DBG): 1 = fwd B tile as 64-row TMA boxes, 2 = skip epilogue stores
  int64_t dw_tap_stride;    // Cout*Cin
  // fwd tail split (wave quantisation): the last, partly filled wave of tiles is cut along K into `tail_parts` pieces per tile;
  // pieces meet in an fp32 scratch tile and the last one to arrive runs the epilogue
  int32_t tail_first;       // index of the first tile of the last wave (== num_tiles: no split)
  int32_t tail_tiles;       // tiles in the last wave
  int32_t tail_parts;       // K pieces per tail tile (<= 1: no split)
  float* scratch;           // [tail_tiles * tail_parts][128][BN] fp32
  int32_t* counters;        // [tail_tiles] arrival counters, zero between launches
  // fwd resident activation slab: rows [m0 + a_row_off, + 128 + (k-1)*dil) of one 64-channel chunk are fetched ONCE and every
  // tap's A operand is a row-shifted view of it (tcgen05 applies the 128B swizzle on absolute shared-memory address bits, so
  // a descriptor may start at any row: tools/probe_umma_row_offset.cu); only the weight tile is fetched per (tap, chunk)
  int32_t slab_rows;        // 0: per-tap A tiles (the K loop is tap-major); > 0: slab mode (the K loop is chunk-major)
  int32_t slab_bytes;       // one slab buffer (slab_rows * 128 rounded up to the 1024-byte swizzle atom); three of them
  int32_t b_stages;         // weight-tile ring depth in slab mode (4, or 3 when the slabs are large)
  int32_t ring_bytes;       // operand rings occupy [0, ring_bytes) of the (1024-aligned) dynamic shared memory; barriers follow
  // CTA-pair kernel (conv_gemm_cg2_kernel): a pair computes 256 x BN -- two M tiles of one utterance (c2_pair_b == 0) or the same
  // M tile of two consecutive utterances (c2_pair_b == 1) -- and each CTA fetches HALF of every weight tile
  int32_t c2_pair_b;
  int32_t c2_mu, c2_bu;     // units along M / along the batch a pair index decodes into
  int32_t c2_ngroup;        // N tiles that run side by side on the same activation rows (their weight slices stay in L2 together)
  int32_t c2_k;             // wgrad pair kernel: the real tap count (p.k then counts tap PAIRS)
  int32_t c2_npad;          // fwd-kind pair kernel: padded column count (the last N tile covers [.., c2_npad))
  // Backward-data GEMM of the layer ABOVE a BatchNorm block (w2l_conv1d_dgrad_wt_bnred): the rows this launch produces are the
  // gradient with respect to that block's padded OUTPUT, so its epilogue also forms the block's backward reduction --
  // r_red[0:C] += sum g, r_red[C:2C] += sum g * (z - mean), g = gate(z) * keep-bit * (this row's gradient, as stored) -- and the
  // separate pass over (dy, z) that w2l_bn_act_bwd_reduce would make is not needed.  A reflect-halo row contributes with the z /
  // gate of the interior row it mirrors, which is exactly the halo fold of that pass.
  const __nv_bfloat16* r_z;     // the block's conv output [r_B, r_T, C], C = N_valid
  const uint8_t* r_mask;        // dropout keep-bits (nullable)
  const float* r_scale;         // the block's BatchNorm scale / shift / mean [C]
  const float* r_shift;
  const float* r_mean;
  const int32_t* r_lens;        // nullable: rows t >= r_lens[b] carry no gradient
  float* r_red;                 // null: no fused reduction
  int32_t r_B, r_T, r_pl, r_Tp, r_act;
  float r_inv_keep;             // 1 without dropout
  // fp32-faithful mode (w2l_conv_desc::x_dtype == F32): operands are fp32 in memory, multiplied as tf32 (kind::tf32, K = 8 per
  // MMA).  A 128-byte swizzle row then holds 32 elements, so a K-step covers kblk = 32 channels (fwd/dgrad) or 32 rows (wgrad);
  // every byte offset of the pipeline (stage sizes, 32 bytes per MMA along K) is the same as with bf16.
  int32_t tf32;
  int32_t kblk;             // elements of the contraction per K-step: 64 (bf16) or 32 (tf32)
  int32_t wg_kmajor;        // wgrad over TRANSPOSED operands dyT [B, Cout, T], xT [B, Cin, rows] (time contiguous): both K-major
  // ... where TMA wants the innermost (time) coordinate 16-byte aligned: tap j reads x at time offset o = off + j*d, so the caller
  // supplies xT DELAYED by s = (-o) mod 4 rows (xT_s[b, ci, u] = x[b, u - s, ci], s zeros in front) and the load uses map tmX[s] at
  // coordinate t0 + o + s, a multiple of 4; coordinates below 0 read as zero, which is the conv's zero padding
  CUtensorMap tmX[4];
};

// A unit of work: (part of) one output tile.  FWD/DGRAD: whole tiles, statically strided over the CTAs; the K loop
// walks (tap, channel chunk).  WGRAD: stream-K -- the global iteration space tiles x (utterance, time chunk) is cut into
// gridDim.x equal contiguous ranges, so every CTA does the same number of MMAs whatever the tile count; a tile that
// straddles two CTAs is accumulated with fp32 atomics (partial), a tile owned by one CTA is stored directly.
struct Unit {
  int m0, n0, b, j;      // b: utterance (fwd/dgrad); j: tap (wgrad)
  int it_begin, it_end;  // K iterations of the tile covered by this unit
  bool partial;
  int slot, tail;        // fwd tail split: scratch slot of this piece, index of the tile within the last wave
};

template <int MODE>
struct UnitIter {
  const GemmParams& p;
  int tile;              // whole-tile cursor (fwd/dgrad; wgrad's data-parallel rounds)
  int tile_end;
  int64_t g, g_end;      // wgrad stream-K cursor in the remainder iteration space
  int iters;             // K iterations per tile
  int cta, ncta;         // this CTA's (or CTA pair's) index among the units of the persistent grid
  __device__ UnitIter(const GemmParams& p_, int cta_ = blockIdx.x, int ncta_ = gridDim.x) : p(p_), cta(cta_), ncta(ncta_) {
    tile = cta;
    if (MODE == MODE_WGRAD) {
      iters = p.B * p.kc_steps;
      tile_end = p.wg_rounds * ncta;
      const int64_t rem = (int64_t)(p.num_tiles - tile_end) * iters;
      g = rem * cta / ncta;
      g_end = rem * (cta + 1) / ncta;
    } else {
      iters = p.k * p.kc_steps;
      tile_end = p.tail_parts > 1 ? p.tail_first : p.num_tiles;
      g = 0;                                      // fwd/dgrad: 0 = the tail piece of this CTA is still to come
    }
  }
  __device__ void decode_wgrad(int t, Unit& u) const {      // tap fastest: neighbouring CTAs share dy and (shifted) x rows
    u.j = t % p.k;
    t /= p.k;
    u.n0 = (t % p.n_tiles) * p.BN;
    u.m0 = (t / p.n_tiles) * kBlockM;
    u.b = 0;
  }
  __device__ void decode_fwd(int t, Unit& u) const {
    const int mt = t % p.m_tiles;
    t /= p.m_tiles;
    u.b = t % p.B;
    u.n0 = (t / p.B) * p.BN;
    u.m0 = mt * kBlockM;
    u.j = 0;
  }
  __device__ bool next(Unit& u) {
    if (tile < tile_end) {
      int t = tile;
      tile += ncta;
      u.it_begin = 0;
      u.it_end = iters;
      u.partial = false;
      if (MODE == MODE_WGRAD) {
        decode_wgrad(t, u);
      } else {
        decode_fwd(t, u);
      }
      return true;
    }
    if (MODE != MODE_WGRAD) {
      if (p.tail_parts <= 1 || g != 0) return false;
      g = 1;
      const int c = cta;
      if (c >= p.tail_tiles * p.tail_parts) return false;
      u.tail = c / p.tail_parts;
      const int part = c - u.tail * p.tail_parts;
      u.slot = c;
      u.it_begin = (int)((int64_t)iters * part / p.tail_parts);
      u.it_end = (int)((int64_t)iters * (part + 1) / p.tail_parts);
      u.partial = true;
      decode_fwd(p.tail_first + u.tail, u);
      return true;
    }
    if (MODE == MODE_WGRAD) {
      if (g >= g_end) return false;
      const int t = (int)(g / iters);
      u.it_begin = (int)(g - (int64_t)t * iters);
      const int64_t left = g_end - g;
      u.it_end = (int)((int64_t)(iters - u.it_begin) <= left ? iters : u.it_begin + left);
      u.partial = (u.it_begin != 0) || (u.it_end != iters);
      g += u.it_end - u.it_begin;
      decode_wgrad(tile_end + t, u);
      return true;
    }
    return false;
  }
};

// Column sums over the 32 rows a warp holds (one row per lane, 32 columns per lane): recursive halving -- at each step a lane
// keeps the half of its columns selected by one bit of its lane id and receives the partner's sums for that half -- 31
// shuffles leave lane i with the sum of column i.
__device__ __forceinline__ float warp_column_sum32(float (&x)[32], int lane) {
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) {
    const bool upper = (lane & w) != 0;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const float keep = upper ? x[i + w] : x[i];
      const float send = upper ? x[i] : x[i + w];
      x[i] = keep + __shfl_xor_sync(0xffffffffu, send, w);
    }
  }
  return x[0];
}

// fwd/dgrad epilogue of one 32-column chunk of a thread's accumulator row: affine (bias / BatchNorm fold) -> activation -> BatchNorm
// batch statistics of what is stored -> bf16 / fp32 store.  `aff` = this chunk's (scale, shift) pairs in shared memory.
template <int MODE>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, float (&v)[32], int c0, int nbase, const float2* aff, bool has_aff,
                                               bool has_act, float act_hi, bool row_ok, int b, int m, int lane,
                                               const float* mean = nullptr) {
  if (MODE == MODE_FWD && p.r_red != nullptr) {      // kernel-uniform: fused BatchNorm-backward reduction of the block below (see GemmParams)
    // (utterance, interior row) of the block's output that this row of the gradient belongs to: M_valid rows per launch batch entry
    // (r_Tp per utterance; the flat launch covers all utterances in one batch entry)
    const int64_t flat = (int64_t)b * p.M_valid + m;
    const int bb = (int)(flat / p.r_Tp);
    int t = (int)(flat - (int64_t)bb * p.r_Tp) - p.r_pl;
    if (t < 0) t = -t;                                 // left halo row: mirrors interior row pl - p
    else if (t >= p.r_T) t = 2 * (p.r_T - 1) - t;      // right halo row
    const bool live = row_ok && bb < p.r_B && !(p.r_lens != nullptr && t >= __ldg(p.r_lens + bb));
    float s1[32], s2[32];
    if (live) {
      const int64_t e = ((int64_t)bb * p.r_T + t) * p.N_valid + nbase;
      const bool full = (c0 + 32 <= p.BN) && (nbase + 32 <= p.N_valid);
      float zf[32];
      if (full) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const uint4 zq = __ldg(reinterpret_cast<const uint4*>(p.r_z + e) + q4);
          const uint32_t w4[4] = {zq.x, zq.y, zq.z, zq.w};                   // two bf16 per word: value = the 16 bits as the top of an fp32
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            zf[q4 * 8 + 2 * i] = __uint_as_float(w4[i] << 16);
            zf[q4 * 8 + 2 * i + 1] = __uint_as_float(w4[i] & 0xFFFF0000u);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) zf[i] = (c0 + i < p.BN && nbase + i < p.N_valid) ? __bfloat162float(p.r_z[e + i]) : 0.f;
      }
      uint32_t bits = 0xFFFFFFFFu;
      if (p.r_mask != nullptr) {
        bits = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (nbase + 8 * q < p.N_valid) bits |= (uint32_t)__ldg(p.r_mask + ((e + 8 * q) >> 3)) << (8 * q);
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const bool col_ok = (c0 + i < p.BN) && (nbase + i < p.N_valid);
        const float2 a = aff[i];                                              // (scale, shift) of the block, 1/keep folded in
        const float pre = fmaf(zf[i], a.x, a.y);
        const bool pass = p.r_act == W2L_ACT_RELU ? pre > 0.f : p.r_act == W2L_ACT_CLAMP20 ? (pre >= 0.f && pre <= 20.f) : true;
        const float gq = __bfloat162float(__float2bfloat16_rn(v[i])) * p.r_inv_keep;   // the gradient as the apply pass will read it
        const float g = (col_ok && pass && ((bits >> i) & 1u)) ? gq : 0.f;
        s1[i] = g;
        s2[i] = g * (zf[i] - mean[i]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) s1[i] = s2[i] = 0.f;
    }
    const float c1 = warp_column_sum32(s1, lane), c2 = warp_column_sum32(s2, lane);
    if (c0 + lane < p.BN && nbase + lane < p.N_valid) {
      atomicAdd(p.r_red + nbase + lane, c1);
      atomicAdd(p.r_red + p.N_valid + nbase + lane, c2);
    }
  }
  if (has_aff) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float2 a = aff[i];
      v[i] = fmaf(v[i], a.x, a.y);
    }
  }
  if (has_act) {   // NaN passes through, as torch.clamp / relu do
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = (v[i] != v[i]) ? v[i] : fminf(fmaxf(v[i], 0.f), act_hi);
  }
  if (MODE == MODE_FWD && p.bn_stats != nullptr) {   // BatchNorm batch statistics of what is stored (kernel-uniform branch)
    float s1[32], s2[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float q = !row_ok ? 0.f : p.y_dtype == W2L_DTYPE_BF16 ? __bfloat162float(__float2bfloat16_rn(v[i])) : v[i];
      s1[i] = q;
      s2[i] = q * q;
    }
    const float c1 = warp_column_sum32(s1, lane), c2 = warp_column_sum32(s2, lane);
    if (c0 + lane < p.BN && nbase + lane < p.N_valid) {
      atomicAdd(p.bn_stats + nbase + lane, c1);
      atomicAdd(p.bn_stats + p.N_valid + nbase + lane, c2);
    }
  }
  if (row_ok && !(p.dbg & 2)) {
    const int64_t off = (int64_t)b * p.y_batch_stride + (int64_t)(m + p.y_row_off) * p.ldy + nbase;
    const bool full = (c0 + 32 <= p.BN) && (nbase + 32 <= p.N_valid);
    if (p.y_dtype == W2L_DTYPE_BF16) {
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.y) + off;
      if (full) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 q;
          q.x = pack_bf16x2(v[i], v[i + 1]);
          q.y = pack_bf16x2(v[i + 2], v[i + 3]);
          q.z = pack_bf16x2(v[i + 4], v[i + 5]);
          q.w = pack_bf16x2(v[i + 6], v[i + 7]);
          *reinterpret_cast<uint4*>(dst + i) = q;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c0 + i < p.BN && nbase + i < p.N_valid) dst[i] = __float2bfloat16_rn(v[i]);
      }
    } else {
      float* dst = reinterpret_cast<float*>(p.y) + off;
      if (full) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c0 + i < p.BN && nbase + i < p.N_valid) dst[i] = v[i];
      }
    }
  }
}

// per-tile epilogue constants into shared memory (the epilogue warps, ep_tid = 0..127; a bar.sync of the 128 follows): the fused
// affine terms of a forward launch, or -- backward-data with the fused BatchNorm-backward reduction -- the block's scale / shift with
// the dropout factor folded in, and its mean
__device__ __forceinline__ void stage_epilogue_constants(const GemmParams& p, float2* aff, float* mean, int n0, int ep_tid, bool has_aff) {
  for (int c = ep_tid; c < p.BN; c += 128) {
    const int n = min(n0 + c, p.N_valid - 1);
    if (has_aff) {
      const float sc = p.scale ? __ldg(p.scale + n) : 1.f;
      const float sh = (p.bias ? __ldg(p.bias + n) * sc : 0.f) + (p.shift ? __ldg(p.shift + n) : 0.f);
      aff[c] = make_float2(sc, sh);
    } else {
      aff[c] = make_float2(__ldg(p.r_scale + n) * p.r_inv_keep, __ldg(p.r_shift + n) * p.r_inv_keep);
      mean[c] = __ldg(p.r_mean + n);
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(kGemmThreads, 1) conv_gemm_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.ring_bytes);
  uint64_t* full_bar = bars;                  // [kStages]
  uint64_t* empty_bar = bars + kStages;       // [kStages]
  uint64_t* tfull_bar = bars + 2 * kStages;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;       // [2]
  uint64_t* slab_full = tempty_bar + 2;       // [kSlabBufs]  slab mode
  uint64_t* slab_empty = slab_full + kSlabBufs;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(slab_empty + kSlabBufs);
  uint32_t* s_flag = tmem_slot + 1;           // fwd tail split: arrival order of this CTA's piece
  // slab mode smem map: [kSlabBufs slabs of slab_bytes][b_stages weight tiles of 32 KB]
  const bool slab_mode = (MODE == MODE_FWD) && p.slab_rows > 0;
  const uint32_t slab_bytes = (uint32_t)p.slab_bytes, b_ring_off = kSlabBufs * slab_bytes;
  float2* s_aff = reinterpret_cast<float2*>(smem + p.ring_bytes + 256);   // [2][256]
  float* s_mean = reinterpret_cast<float*>(s_aff + 2 * 256);               // [2][256]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr bool kAMN = (MODE == MODE_WGRAD);
  constexpr bool kBMN = (MODE != MODE_FWD);
  const int b_chunks = (p.BN + 63) >> 6;
  const bool b_mn = kBMN && !p.wg_kmajor, a_mn = kAMN && !p.wg_kmajor;      // kernel-uniform
  const uint32_t stage_tx = kABytes + ((b_mn || (p.dbg & 1)) ? (uint32_t)b_chunks * 8192u : (uint32_t)p.BN * 128u);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 128);
    }
    for (int i = 0; i < kSlabBufs; ++i) {
      mbar_init(&slab_full[i], 1);
      mbar_init(&slab_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 2 * kAccCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer
      int stage = 0;
      uint32_t phase = 0;
      UnitIter<MODE> units(p);
      Unit u;
      if (slab_mode) {
        int sidx = 0;
        uint32_t sphase = 0;
        while (units.next(u)) {
          const int kc_last = (u.it_end - 1) / p.k;
          int kc_issued = u.it_begin / p.k;                     // next chunk whose slab has not been requested yet
          for (int it = u.it_begin; it < u.it_end; ++it) {
            const int kc = it / p.k, j = it - kc * p.k;
            // keep the slab of the NEXT chunk in flight while this chunk's weight tiles stream: with three buffers the one it
            // lands in was released a whole chunk ago, so this never waits
            while (kc_issued <= kc_last && kc_issued <= kc + 1) {
              mbar_wait(&slab_empty[sidx], sphase ^ 1u);
              mbar_expect_tx(&slab_full[sidx], (uint32_t)p.slab_rows * 128u);
              tma_load_3d(smem + sidx * slab_bytes, &p.tmA, &slab_full[sidx], kc_issued * kBlockK, u.m0 + p.a_row_off, u.b);
              ++kc_issued;
              if (++sidx == kSlabBufs) {
                sidx = 0;
                sphase ^= 1u;
              }
            }
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            mbar_expect_tx(&full_bar[stage], (uint32_t)p.BN * 128u);
            tma_load_3d(smem + b_ring_off + stage * kBBytesMax, &p.tmB, &full_bar[stage], kc * kBlockK, u.n0, j);
            if (++stage == p.b_stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      } else
      while (units.next(u)) {
        for (int it = u.it_begin; it < u.it_end; ++it) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_expect_tx(&full_bar[stage], stage_tx);
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + kABytes;
          if (MODE == MODE_WGRAD) {
            const int b = it / p.kc_steps, t0 = (it - b * p.kc_steps) * p.kblk;
            const int xr = t0 + p.b_row_off + u.j * p.dil;
            if (p.wg_kmajor) {  // transposed operands: the contraction (time) is the innermost coordinate of both maps
              const int o = p.b_row_off + u.j * p.dil, sft = (-o) & 3;       // two's complement: o + sft is a multiple of 4 for o < 0 too
              tma_load_3d(sa, &p.tmA, &full_bar[stage], t0, u.m0, b);
              tma_load_3d(sb, &p.tmX[sft], &full_bar[stage], t0 + o + sft, u.n0, b);
            } else if (p.mn4d) {       // one TMA op per operand: box (64 ch, 64 rows, chunks, 1) lands as [chunk][row][64]
              tma_load_4d(sa, &p.tmA, &full_bar[stage], 0, t0, u.m0 >> 6, b);
              tma_load_4d(sb, &p.tmB, &full_bar[stage], 0, xr, u.n0 >> 6, b);
            } else {
              tma_load_3d(sa, &p.tmA, &full_bar[stage], u.m0, t0, b);
              tma_load_3d(sa + 8192, &p.tmA, &full_bar[stage], u.m0 + 64, t0, b);
              for (int c = 0; c < b_chunks; ++c) tma_load_3d(sb + c * 8192, &p.tmB, &full_bar[stage], u.n0 + c * 64, xr, b);
            }
          } else {
            const int j = it / p.kc_steps, kc = it - j * p.kc_steps;
            tma_load_3d(sa, &p.tmA, &full_bar[stage], kc * p.kblk, u.m0 + p.a_row_off + j * p.a_tap_step, u.b);
            if (MODE == MODE_FWD) {
              if (p.dbg & 1) {
                for (int c = 0; c < b_chunks; ++c) tma_load_3d(sb + c * 8192, &p.tmB, &full_bar[stage], kc * p.kblk, u.n0 + c * 64, j);
              } else {
                tma_load_3d(sb, &p.tmB, &full_bar[stage], kc * p.kblk, u.n0, j);
              }
            } else {
              for (int c = 0; c < b_chunks; ++c)
                tma_load_3d(sb + c * 8192, &p.tmB, &full_bar[stage], u.n0 + c * 64, kc * kBlockK, j);
            }
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------ MMA issuer
      const uint32_t idesc = make_idesc_bf16(kBlockM, p.BN, a_mn, b_mn, p.tf32 != 0);
      const uint32_t a_kstep = a_mn ? 2048u : 32u;       // bytes per UMMA_K (16 bf16 / 8 tf32 elements of K)
      const uint32_t b_kstep = b_mn ? 2048u : 32u;
      const uint32_t a_lbo = a_mn ? 8192u : 16u;
      const uint32_t b_lbo = b_mn ? 8192u : 16u;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      UnitIter<MODE> units(p);
      Unit u;
      if (slab_mode) {
        int sidx = 0;
        uint32_t sphase = 0;
        while (units.next(u)) {
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)acc * kAccCols;
          uint32_t accumulate = 0;
          int cur_kc = -1, cur_s = 0;
          uint32_t a_base = 0;
          for (int it = u.it_begin; it < u.it_end; ++it) {
            const int kc = it / p.k, j = it - kc * p.k;
            if (kc != cur_kc) {
              if (cur_kc >= 0) umma_commit(&slab_empty[cur_s]);   // every MMA that read the previous slab has been issued
              cur_kc = kc;
              cur_s = sidx;
              mbar_wait(&slab_full[sidx], sphase);
              tc_fence_after();
              a_base = smem_u32(smem + sidx * slab_bytes);
              if (++sidx == kSlabBufs) {
                sidx = 0;
                sphase ^= 1u;
              }
            }
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t a_addr = a_base + (uint32_t)(j * p.dil) * 128u;      // tap j = the slab shifted by j*dilation rows
            const uint32_t b_addr = smem_u32(smem + b_ring_off + stage * kBBytesMax);
#pragma unroll
            for (int kk = 0; kk < kBlockK / 16; ++kk) {
              const uint64_t adesc = make_smem_desc(a_addr + kk * 32u, 16u, 1024u);
              const uint64_t bdesc = make_smem_desc(b_addr + kk * 32u, 16u, 1024u);
              umma_bf16(d_tmem, adesc, bdesc, idesc, accumulate);
              accumulate = 1;
            }
            umma_commit(&empty_bar[stage]);
            if (++stage == p.b_stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
          umma_commit(&slab_empty[cur_s]);
          umma_commit(&tfull_bar[acc]);
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1u;
        }
      } else
      while (units.next(u)) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * kAccCols;
        uint32_t accumulate = 0;
        for (int it = u.it_begin; it < u.it_end; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * kStageBytes);
          const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
          for (int kk = 0; kk < kBlockK / 16; ++kk) {
            const uint64_t adesc = make_smem_desc(a_addr + kk * a_kstep, a_lbo, 1024u);
            const uint64_t bdesc = make_smem_desc(b_addr + kk * b_kstep, b_lbo, 1024u);
            if (p.tf32)
              umma_tf32(d_tmem, adesc, bdesc, idesc, accumulate);
            else
              umma_bf16(d_tmem, adesc, bdesc, idesc, accumulate);
            accumulate = 1;
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // -------------------------------------------------------------- epilogue (warps 2..5)
    const int lane_base = (warp & 3) * 32;
    const int row = lane_base + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    // fused epilogue: v = clamp(acc * scale + shift', lo, hi) with shift' = bias * scale + shift (kernel-uniform switches)
    const bool has_aff = (MODE == MODE_FWD) && (p.bias != nullptr || p.scale != nullptr);
    const bool has_red = (MODE == MODE_FWD) && p.r_red != nullptr;      // (never together with has_aff: forward vs backward-data launches)
    const bool has_act = (MODE == MODE_FWD) && p.act != W2L_ACT_NONE;
    const float act_hi = p.act == W2L_ACT_CLAMP20 ? 20.f : INFINITY;
    const int ep_tid = threadIdx.x - 64;                       // 0..127 within the epilogue warps
    UnitIter<MODE> units(p);
    Unit tc;
    while (units.next(tc)) {
      if (has_aff || has_red) {                                // stage this tile's per-channel constants (broadcast reads below)
        stage_epilogue_constants(p, s_aff + acc * 256, s_mean + acc * 256, tc.n0, ep_tid, has_aff);
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)acc * kAccCols;
      const int m = tc.m0 + row;
      const bool row_ok = m < p.M_valid;
      auto finish_chunk = [&](float (&v)[32], int c0, int nbase) {
        epilogue_chunk<MODE>(p, v, c0, nbase, s_aff + acc * 256 + c0, has_aff, has_act, act_hi, row_ok, tc.b, m, lane,
                             s_mean + acc * 256 + c0);
      };
      if (MODE != MODE_WGRAD && tc.partial) {
        // ---- fwd tail split: this CTA computed one K piece of the tile.  Park the fp32 partial in the scratch slot, free the
        // accumulator, and let the LAST piece to arrive add the pieces in a fixed order and run the epilogue.
        float* mine = p.scratch + ((int64_t)tc.slot * kBlockM + row) * p.BN;
        for (int c0 = 0; c0 < p.BN; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(t_addr + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            if (c0 + i < p.BN)
              *reinterpret_cast<float4*>(mine + c0 + i) = make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                                                                      __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
        }
        tc_fence_before();
        mbar_arrive(&tempty_bar[acc]);
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (ep_tid == 0) *s_flag = (uint32_t)atomicAdd(p.counters + tc.tail, 1);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if ((int)*s_flag == p.tail_parts - 1) {
          __threadfence();
          const float* first = p.scratch + ((int64_t)tc.tail * p.tail_parts * kBlockM + row) * p.BN;
          for (int c0 = 0; c0 < p.BN; c0 += 32) {
            const int nbase = tc.n0 + c0;
            if (nbase >= p.N_valid) break;
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
            for (int q = 0; q < p.tail_parts; ++q) {
              const float* src = first + (int64_t)q * kBlockM * p.BN + c0;
#pragma unroll
              for (int i = 0; i < 32; i += 4)
                if (c0 + i < p.BN) {
                  const float4 t4 = __ldcg(reinterpret_cast<const float4*>(src + i));
                  v[i] += t4.x;
                  v[i + 1] += t4.y;
                  v[i + 2] += t4.z;
                  v[i + 3] += t4.w;
                }
            }
            finish_chunk(v, c0, nbase);
          }
          if (ep_tid == 0) p.counters[tc.tail] = 0;          // clean for the next launch
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");          // s_flag is reused by the next unit
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
        continue;
      }
      for (int c0 = 0; c0 < p.BN; c0 += 32) {
        const int nbase = tc.n0 + c0;
        if (nbase >= p.N_valid) break;                       // padded part of a last N tile
        uint32_t r[32];
        tmem_ld_32x32(t_addr + c0, r);
        tmem_ld_wait();
        if (MODE == MODE_WGRAD) {
          if (row_ok) {
            float* dst = reinterpret_cast<float*>(p.y) + (int64_t)tc.j * p.dw_tap_stride + (int64_t)m * p.ldy + nbase;
            if (tc.partial) {
              if (c0 + 32 <= p.BN && nbase + 32 <= p.N_valid) {
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "f"(__uint_as_float(r[i])),
                               "f"(__uint_as_float(r[i + 1])), "f"(__uint_as_float(r[i + 2])), "f"(__uint_as_float(r[i + 3]))
                               : "memory");
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (c0 + i < p.BN && nbase + i < p.N_valid) atomicAdd(dst + i, __uint_as_float(r[i]));
              }
            } else if (c0 + 32 <= p.BN && nbase + 32 <= p.N_valid) {
#pragma unroll
              for (int i = 0; i < 32; i += 4)
                *reinterpret_cast<float4*>(dst + i) = make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                                                                  __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (c0 + i < p.BN && nbase + i < p.N_valid) dst[i] = __uint_as_float(r[i]);
            }
          }
        } else {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          finish_chunk(v, c0, nbase);
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * kAccCols);
  }
}


// ---------------------------------------------------------------- CTA-pair variant (cta_group::2), FWD-kind operands (both K-major)
// Same roles as conv_gemm_kernel, but the two CTAs of a cluster share every weight tile: each fetches A for its own 128 rows and
// HALF of the B tile (BN/2 weight rows); the leader's single thread issues tcgen05.mma.cta_group::2 (M = 256), which reads both
// CTAs' shared memory and fills a 128 x BN accumulator in EACH CTA's TMEM.  Per CTA and K-step that is 16 KB + BN*64 B from L2
// instead of 16 KB + BN*128 B: the single-CTA kernel needs ~17 TB/s of L2->SM traffic at 1.4 PFLOP/s (83 FLOP/B at BN = 224),
// this one 122 FLOP/B, and the ring holds 6 K-steps instead of 4.
//   full[s]   (leader)  : 1 arrival (leader's producer, expect_tx of BOTH CTAs' bytes) + the bytes of both CTAs' TMA loads
//   empty[s]  (each CTA): tcgen05.commit multicast to both CTAs -> each producer re-fills its own stage
//   tfull[a]  (each CTA): tcgen05.commit multicast -> each CTA's epilogue warps read their own TMEM
//   tempty[a] (leader)  : 8 arrivals = one per epilogue warp of both CTAs (remote arrive from the peer)
constexpr int kC2Stages = 6;
constexpr int kC2BBytes = 128 * kBlockK * 2;            // half of a <= 256-wide weight tile
constexpr int kC2StageBytes = kABytes + kC2BBytes;      // 32 KB
constexpr size_t kC2Smem = (size_t)kC2Stages * kC2StageBytes + 1024 + 256 + kAffBytes;

struct PairTile {
  int m0, n0, b;
};
__device__ __forceinline__ void decode_pair(const GemmParams& p, int t, int rank, PairTile& u) {
  const int ni = t % p.c2_ngroup;
  t /= p.c2_ngroup;
  const int mu = t % p.c2_mu;
  t /= p.c2_mu;
  const int bu = t % p.c2_bu;
  const int grp = t / p.c2_bu;
  u.n0 = (grp * p.c2_ngroup + ni) * p.BN;
  if (p.c2_pair_b) {
    u.m0 = mu * kBlockM;
    u.b = bu * 2 + rank;
  } else {
    u.m0 = (mu * 2 + rank) * kBlockM;
    u.b = bu;
  }
}

// width of the N tile that starts at column n0: BN, except that the LAST tile of a row ends at the padded column count c2_npad (a
// multiple of 16; 640 columns run as 224 + 224 + 192).  The TMA box stays BN/2 weight rows per CTA; a narrow tile uses the first
// rows of what each CTA fetched.
__device__ __forceinline__ int c2_tile_bn(const GemmParams& p, int n0) {
  const int left = p.c2_npad - n0;
  return left < p.BN ? left : p.BN;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1) conv_gemm_cg2_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);     // same offset in both CTAs (same kernel, same static layout)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kC2Stages * kC2StageBytes);
  uint64_t* full_bar = bars;                      // [kC2Stages]
  uint64_t* empty_bar = bars + kC2Stages;         // [kC2Stages]
  uint64_t* tfull_bar = bars + 2 * kC2Stages;     // [2]
  uint64_t* tempty_bar = tfull_bar + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float2* s_aff = reinterpret_cast<float2*>(smem + kC2Stages * kC2StageBytes + 256);   // [2][256]
  float* s_mean = reinterpret_cast<float*>(s_aff + 2 * 256);                            // [2][256]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int iters = p.k * p.kc_steps;
  const uint32_t half_bn = (uint32_t)p.BN >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    for (int i = 0; i < kC2Stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_cg2(tmem_slot, 2 * kAccCols);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  cluster_sync_all();                             // barriers of BOTH CTAs are initialised before anyone signals across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer (both CTAs; bytes land on the LEADER's barrier)
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t stage_tx = 2u * ((uint32_t)kABytes + half_bn * 128u);
      PairTile u;
      for (int t = pair; t < p.num_tiles; t += npairs) {
        decode_pair(p, t, (int)rank, u);
        const int half_u = c2_tile_bn(p, u.n0) >> 1;
        for (int it = 0; it < iters; ++it) {
          const int j = it / p.kc_steps, kc = it - j * p.kc_steps;
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          if (rank == 0) mbar_expect_tx(&full_bar[stage], stage_tx);
          const uint32_t lead_full = mapa_shared(smem_u32(&full_bar[stage]), 0);
          uint8_t* sa = smem + stage * kC2StageBytes;
          tma_load_3d_cg2(sa, &p.tmA, lead_full, kc * kBlockK, u.m0 + p.a_row_off + j * p.a_tap_step, u.b);
          tma_load_3d_cg2(sa + kABytes, &p.tmB, lead_full, kc * kBlockK, u.n0 + (int)rank * half_u, j);
          if (++stage == kC2Stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ------------------------------------------------------------ MMA issuer (leader CTA only)
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      PairTile u;
      for (int t = pair; t < p.num_tiles; t += npairs) {
        decode_pair(p, t, 0, u);
        const uint32_t idesc = make_idesc_bf16(2 * kBlockM, c2_tile_bn(p, u.n0), false, false);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * kAccCols;
        uint32_t accumulate = 0;
        for (int it = 0; it < iters; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * kC2StageBytes);
          const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
          for (int kk = 0; kk < kBlockK / 16; ++kk) {
            const uint64_t adesc = make_smem_desc(a_addr + kk * 32u, 16u, 1024u);
            const uint64_t bdesc = make_smem_desc(b_addr + kk * 32u, 16u, 1024u);
            umma_bf16_cg2(d_tmem, adesc, bdesc, idesc, accumulate);
            accumulate = 1;
          }
          umma_commit_cg2(&empty_bar[stage], 3);
          if (++stage == kC2Stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit_cg2(&tfull_bar[acc], 3);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // -------------------------------------------------------------- epilogue (warps 2..5 of both CTAs, each on its own TMEM)
    const int lane_base = (warp & 3) * 32;
    const int row = lane_base + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool has_aff = p.bias != nullptr || p.scale != nullptr;
    const bool has_red = p.r_red != nullptr;
    const bool has_act = p.act != W2L_ACT_NONE;
    const float act_hi = p.act == W2L_ACT_CLAMP20 ? 20.f : INFINITY;
    const int ep_tid = threadIdx.x - 64;
    PairTile u;
    for (int t = pair; t < p.num_tiles; t += npairs) {
      decode_pair(p, t, (int)rank, u);
      if (has_aff || has_red) {
        stage_epilogue_constants(p, s_aff + acc * 256, s_mean + acc * 256, u.n0, ep_tid, has_aff);
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)acc * kAccCols;
      const int m = u.m0 + row;
      const bool row_ok = m < p.M_valid && u.b < p.B;
      const int bn_u = c2_tile_bn(p, u.n0);
      for (int c0 = 0; c0 < bn_u; c0 += 32) {
        const int nbase = u.n0 + c0;
        if (nbase >= p.N_valid) break;
        uint32_t r[32];
        tmem_ld_32x32(t_addr + c0, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
        epilogue_chunk<MODE_FWD>(p, v, c0, nbase, s_aff + acc * 256 + c0, has_aff, has_act, act_hi, row_ok, u.b, m, lane,
                                 s_mean + acc * 256 + c0);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[acc]), 0));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }
  tc_fence_before();
  cluster_sync_all();                             // the peer's shared memory and barriers stay alive until every MMA has retired
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, 2 * kAccCols);
  }
}

// ---------------------------------------------------------------- CTA-pair weight gradient (cta_group::2)
// Transposed problem, so that the operand the two CTAs SHARE is the one that does not depend on the tap:
//   dW_j^T[ci, co] = sum_{b,t} X[b, t + off + j*d, ci] * dY[b, t, co]        A = X (MN-major, shifted per tap)   B = dY (MN-major)
// The pair works on two consecutive (ci tile, tap) combinations -- tap fastest -- of one BN-wide co tile: each CTA fetches ITS X
// rows and half of the dY tile.
// Work units are cut stream-K over the PAIRS (UnitIter with pair index / pair count); a tile shared between pairs is accumulated
// with fp32 atomics.  The accumulator row is ci, its columns are co: column i of a warp's 32 rows is 128 contiguous bytes of
// dw[j][co][:], so the scalar stores / reductions below are coalesced.
// width of the N tile that starts at column n0: the last tile of a row is cut to the next multiple of 128 (896 = 3 x 256 + 128, no
// padded MMA work); the TMA box stays BN/2 columns per CTA -- a narrow tile just uses the first chunk of what each CTA fetched
__device__ __forceinline__ int wg2_tile_bn(const GemmParams& p, int n0) {
  const int left = (p.N_valid - n0 + 127) & ~127;
  return left < p.BN ? left : p.BN;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1) conv_wgrad_cg2_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kC2Stages * kC2StageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kC2Stages;
  uint64_t* tfull_bar = bars + 2 * kC2Stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const uint32_t half_bn = (uint32_t)p.BN >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    for (int i = 0; i < kC2Stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_cg2(tmem_slot, 2 * kAccCols);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer (both CTAs)
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t stage_tx = 2u * ((uint32_t)kABytes + half_bn * 128u);
      UnitIter<MODE_WGRAD> units(p, pair, npairs);
      Unit u;
      while (units.next(u)) {
        const int q = u.j * 2 + (int)rank, mt = q / p.c2_k, tap = q - mt * p.c2_k;      // this CTA's (ci tile, tap); past the end: all zero
        const int half_u = wg2_tile_bn(p, u.n0) >> 1;
        for (int it = u.it_begin; it < u.it_end; ++it) {
          const int b = it / p.kc_steps, t0 = (it - b * p.kc_steps) * kBlockK;
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          if (rank == 0) mbar_expect_tx(&full_bar[stage], stage_tx);
          const uint32_t lead_full = mapa_shared(smem_u32(&full_bar[stage]), 0);
          uint8_t* sa = smem + stage * kC2StageBytes;
          tma_load_4d_cg2(sa, &p.tmA, lead_full, 0, t0 + p.b_row_off + tap * p.dil, mt * 2, b);
          tma_load_4d_cg2(sa + kABytes, &p.tmB, lead_full, 0, t0, (u.n0 + (int)rank * half_u) >> 6, b);
          if (++stage == kC2Stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ------------------------------------------------------------ MMA issuer (leader CTA only)
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      UnitIter<MODE_WGRAD> units(p, pair, npairs);
      Unit u;
      while (units.next(u)) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * kAccCols;
        const uint32_t idesc = make_idesc_bf16(2 * kBlockM, wg2_tile_bn(p, u.n0), true, true);
        uint32_t accumulate = 0;
        for (int it = u.it_begin; it < u.it_end; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * kC2StageBytes);
          const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
          for (int kk = 0; kk < kBlockK / 16; ++kk) {
            const uint64_t adesc = make_smem_desc(a_addr + kk * 2048u, 8192u, 1024u);
            const uint64_t bdesc = make_smem_desc(b_addr + kk * 2048u, 8192u, 1024u);
            umma_bf16_cg2(d_tmem, adesc, bdesc, idesc, accumulate);
            accumulate = 1;
          }
          umma_commit_cg2(&empty_bar[stage], 3);
          if (++stage == kC2Stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit_cg2(&tfull_bar[acc], 3);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // -------------------------------------------------------------- epilogue (warps 2..5 of both CTAs)
    const int lane_base = (warp & 3) * 32;
    int acc = 0;
    uint32_t acc_phase = 0;
    UnitIter<MODE_WGRAD> units(p, pair, npairs);
    Unit u;
    while (units.next(u)) {
      const int q = u.j * 2 + (int)rank, mt = q / p.c2_k, tap = q - mt * p.c2_k;
      const int bn_u = wg2_tile_bn(p, u.n0);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)acc * kAccCols;
      const int ci = mt * kBlockM + lane_base + lane;
      const bool row_ok = ci < p.M_valid;
      float* dst0 = reinterpret_cast<float*>(p.y) + (int64_t)tap * p.dw_tap_stride + ci;
      for (int c0 = 0; c0 < bn_u; c0 += 32) {
        const int nbase = u.n0 + c0;
        if (nbase >= p.N_valid) break;
        uint32_t r[32];
        tmem_ld_32x32(t_addr + c0, r);
        tmem_ld_wait();
        if (row_ok) {
          float* dst = dst0 + (int64_t)nbase * p.ldy;
          if (u.partial) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (nbase + i < p.N_valid) atomicAdd(dst + (int64_t)i * p.ldy, __uint_as_float(r[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (nbase + i < p.N_valid) dst[(int64_t)i * p.ldy] = __uint_as_float(r[i]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[acc]), 0));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, 2 * kAccCols);
  }
}

// ---------------------------------------------------------------- host side
// N-tile width: a multiple of `multiple` (16 = UMMA granularity; 64 when operands come through 4-D chunked TMA maps).
// Narrow tiles run the tensor pipe below peak (the A tile is re-read from shared memory per MMA whatever N is; measured
// 1.55 / 1.40 / 1.11 PFLOP/s at N = 256 / 224 / 160), so a wider tile with a partly empty last tile can win: minimise
// (tiles * bn) / eff(bn) with the linear fit eff = 1 - 0.75 * (256 - bn) / 256.
static int pick_bn(int n, int multiple) {
  const int n_round = (n + multiple - 1) / multiple * multiple;
  int best = multiple;
  double best_cost = 1e30;
  for (int bn = multiple; bn <= 256; bn += multiple) {
    if (bn > n_round) break;
    const int tiles = (n + bn - 1) / bn;
    const double eff = 1.0 - 0.75 * (256 - bn) / 256.0;
    const double cost = (double)tiles * bn / eff;
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

// rows of the resident activation slab for a (k, dilation) conv, or 0 when slab mode does not apply.  OPT-IN (W2L_SLAB=1):
// measured on B200 it lowers power (SM clock under the cap 1.53 -> 1.66 GHz: a third less L2->smem traffic) but costs ~17 % more
// tensor-pipe cycles -- a tap's view starts at a row that is not a multiple of the 8-row swizzle atom, and the operand fetch of
// such a view is slower -- so the step is 43.3 ms instead of 40.8 (tools/ab.sh W2L_SLAB 0 1).  Kept as a validated experiment
// (all parity tests pass with it on); the per-tap A tiles stay the default.
static int slab_rows_for(int k, int dil) {
  static const bool enabled = getenv("W2L_SLAB") && atoi(getenv("W2L_SLAB")) == 1;
  const int rows = kBlockM + (k - 1) * dil;
  return (enabled && rows <= 256) ? rows : 0;          // one TMA box holds at most 256 rows
}

// operand-ring geometry of a launch: slab mode (three slabs + a weight-tile ring) when it applies and fits, else the 4-stage A+B ring
static void plan_rings(GemmParams& p, int k, int dil, bool allow_slab) {
  p.ring_bytes = kStages * kStageBytes;
  p.slab_rows = 0;
  const int rows = allow_slab ? slab_rows_for(k, dil) : 0;
  if (!rows) return;
  const int slab_bytes = (rows * 128 + 1023) / 1024 * 1024;
  for (int bs = kStages; bs >= 3; --bs) {
    const size_t ring = (size_t)kSlabBufs * slab_bytes + (size_t)bs * kBBytesMax;
    if (ring + 1024 + 256 + kAffBytes <= kGemmSmemMax) {
      p.slab_rows = rows;
      p.slab_bytes = slab_bytes;
      p.b_stages = bs;
      p.ring_bytes = (int32_t)ring;
      return;
    }
  }
}

// fp32 scratch for the forward tail split, registered by the host (w2l_set_gemm_scratch): [4096 bytes of arrival counters][slots]
static void* g_scratch = nullptr;
static size_t g_scratch_bytes = 0;
constexpr size_t kScratchCounterBytes = 4096;

// Cut the last, partly filled wave of forward tiles along K when that shortens it (see GemmParams::tail_*).
static void plan_tail_split(GemmParams& p) {
  const int grid = p.num_tiles < gemm_sms() ? p.num_tiles : gemm_sms();
  const int iters = p.k * p.kc_steps;
  static const bool enabled = !(getenv("W2L_FWD_TAIL_SPLIT") && atoi(getenv("W2L_FWD_TAIL_SPLIT")) == 0);
  p.tail_first = p.num_tiles;
  if (!enabled || !g_scratch || grid < 2 || p.num_tiles <= grid || iters < 32) return;
  const int full = (p.num_tiles / grid) * grid, rem = p.num_tiles - full;
  if (rem == 0 || rem * 4 > grid * 3) return;                       // the last wave is (nearly) full anyway
  int parts = grid / rem;
  if (parts > 8) parts = 8;
  while (parts > 1 && iters / parts < 8) --parts;
  const size_t need = kScratchCounterBytes + (size_t)rem * parts * kBlockM * p.BN * sizeof(float);
  if (parts < 2 || need > g_scratch_bytes || (size_t)rem * sizeof(int32_t) > kScratchCounterBytes) return;
  p.tail_first = full;
  p.tail_tiles = rem;
  p.tail_parts = parts;
  p.counters = reinterpret_cast<int32_t*>(g_scratch);
  p.scratch = reinterpret_cast<float*>(reinterpret_cast<char*>(g_scratch) + kScratchCounterBytes);
}

template <int MODE>
static int launch_gemm(const GemmParams& p_in, cudaStream_t st, int grid_override = 0) {
  GemmParams p = p_in;
  if (p.ring_bytes == 0) p.ring_bytes = kStages * kStageBytes;      // launchers that do not plan rings: the A+B stage ring
  if (p.kblk == 0) p.kblk = kBlockK;                                // bf16 launchers that never set the operand type
  static bool configured = false;
  if (!configured) {
    W2L_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmemMax));
    configured = true;
  }
  int grid = grid_override > 0 ? grid_override : (p.num_tiles < gemm_sms() ? p.num_tiles : gemm_sms());
  if (grid < 1) return W2L_OK;
  const size_t smem_bytes = (size_t)p.ring_bytes + 1024 /*align*/ + 256 /*barriers*/ + kAffBytes;
  conv_gemm_kernel<MODE><<<grid, kGemmThreads, smem_bytes, st>>>(p);
  return after_launch(MODE == MODE_FWD ? "conv_gemm_kernel<fwd>" : MODE == MODE_DGRAD ? "conv_gemm_kernel<dgrad>" : "conv_gemm_kernel<wgrad>");
}

// CTA-pair kernel: on by default for the FWD-kind GEMMs (forward, backward-data with the transposed weight shadow);
// W2L_CG2=0 falls back to the single-CTA kernel (the A/B switch of profiles/r2_gemm_cg2.md)
// N tile of the fwd-kind pair kernel: equal tiles from pick_bn (the last one ends at the padded column count).  Measured against
// 256-wide tiles with a narrow last one (896 = 3 x 256 + 128): 34.1 / 33.6 vs 34.4 / 34.2 ms of conv time per step -- K-major tiles
// cost in proportion to their width, so fewer-but-wider buys nothing (profiles/r2_gemm_cg2.md).
static int cg2_bn(int npad) { return pick_bn(npad, 16); }

static bool cg2_wanted() {
  const char* e = getenv("W2L_CG2");          // read per call: the parity tests run every case under both kernels in one process
  return !(e && atoi(e) == 0);
}

// geometry of the pair decomposition for a FWD-kind problem whose p.B / p.m_tiles / p.n_tiles / p.BN / p.k / kc_steps are set
static void plan_cg2(GemmParams& p, int64_t weight_bytes_per_ntile) {
  // pair along M when the M tiles pair up (or nothing else does); along the batch when that wastes less
  const int waste_m = p.m_tiles & 1, waste_b = p.B & 1;
  p.c2_pair_b = (waste_m && (!waste_b || p.B > p.m_tiles)) ? 1 : 0;
  p.c2_mu = p.c2_pair_b ? p.m_tiles : (p.m_tiles + 1) / 2;
  p.c2_bu = p.c2_pair_b ? (p.B + 1) / 2 : p.B;
  // N tiles of one group run on the same activation rows at the same time (activations cross DRAM once per group instead of once
  // per N tile) as long as the group's weight slices fit in L2 beside the streams: <= 24 MB
  int g = 1;
  while (g * 2 <= p.n_tiles && p.n_tiles % (g * 2) == 0 && (int64_t)(g * 2) * weight_bytes_per_ntile <= (24ll << 20)) g *= 2;
  p.c2_ngroup = g;
  p.num_tiles = p.c2_mu * p.c2_bu * p.n_tiles;
}

static int launch_gemm_cg2(const GemmParams& p, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    W2L_CUDA(cudaFuncSetAttribute(conv_gemm_cg2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kC2Smem));
    configured = true;
  }
  const int pairs_max = gemm_sms() / 2;
  const int pairs = p.num_tiles < pairs_max ? p.num_tiles : pairs_max;
  if (pairs < 1) return W2L_OK;
  conv_gemm_cg2_kernel<<<2 * pairs, kGemmThreads, kC2Smem, st>>>(p);
  return after_launch("conv_gemm_cg2_kernel");
}

// operand type of a launch from the descriptor: bf16 (default) or fp32 storage multiplied as tf32
static void set_operand_type(GemmParams& p, const w2l_conv_desc* d) {
  p.tf32 = d->x_dtype == W2L_DTYPE_F32;
  p.kblk = p.tf32 ? 32 : kBlockK;
}

static int check_desc(const w2l_conv_desc* d, const char* who) {
  W2L_REQUIRE(d != nullptr, "%s: null descriptor", who);
  W2L_REQUIRE(d->B >= 1 && d->T_out >= 1 && d->k >= 1 && d->dilation >= 1, "%s: bad B/T_out/k/dilation", who);
  W2L_REQUIRE(d->Cin >= 64 && d->Cin % 8 == 0, "%s: Cin=%d must be >= 64 and a multiple of 8", who, d->Cin);
  W2L_REQUIRE(d->Cout >= 1 && d->Cout_pad >= d->Cout && d->Cout_pad % 16 == 0, "%s: Cout=%d Cout_pad=%d (pad must be a multiple of 16)",
              who, d->Cout, d->Cout_pad);
  W2L_REQUIRE(d->x_rows >= 1 && d->y_rows >= d->T_out + d->y_row_offset && d->y_row_offset >= 0, "%s: bad row geometry", who);
  W2L_REQUIRE(d->ldy % 8 == 0, "%s: ldy=%d must be a multiple of 8", who, d->ldy);
  return W2L_OK;
}

// 4-D chunked maps need whole 64-channel chunks (a chunk that ran past the row end would alias the next row)
static bool wgrad_mn4d(const w2l_conv_desc* d) {
  return (d->ldy % 64 == 0) && (d->Cin % 64 == 0) && (d->Cout == d->ldy || d->Cout_pad == d->ldy);
}

// Launch plan for wgrad: `rounds` whole tiles per CTA (data parallel, all CTAs sweep (b, t) in step -> L2 reuse), then the
// remaining tiles*iters iterations are cut into equal contiguous ranges (stream-K).  zero != 0 => dw must be pre-zeroed.
static void wgrad_plan(const w2l_conv_desc* d, int* bn_out, int* grid_out, int* rounds_out, int* zero_out) {
  const int n_pad = (d->Cin + 15) / 16 * 16;
  const int bn = pick_bn(n_pad, wgrad_mn4d(d) ? 64 : 16);
  const int64_t tiles = (int64_t)d->k * ((d->Cout + kBlockM - 1) / kBlockM) * ((n_pad + bn - 1) / bn);
  const int64_t iters = (int64_t)d->B * ((d->T_out + kBlockK - 1) / kBlockK);
  int64_t grid = gemm_sms();
  const int64_t min_iters = 32;                       // keep the per-CTA mainloop long enough to amortise the epilogue
  if (tiles * iters / grid < min_iters) grid = tiles * iters / min_iters > 0 ? tiles * iters / min_iters : 1;
  const int64_t rounds = tiles / grid;
  const int64_t rem_tiles = tiles - rounds * grid;
  if (bn_out) *bn_out = bn;
  if (grid_out) *grid_out = (int)grid;
  if (rounds_out) *rounds_out = (int)rounds;
  if (zero_out) *zero_out = rem_tiles > 0;
}

// ---- CTA-pair wgrad: eligibility and plan.  Needs whole 64-channel chunks on both operands (4-D chunked maps) and at least two taps.
static bool wgrad_cg2_ok(const w2l_conv_desc* d) {
  const char* e = getenv("W2L_CG2_WGRAD");    // 0: single-CTA wgrad beside the pair fwd/dgrad kernels (A/B switch)
  return cg2_wanted() && !(e && atoi(e) == 0) && wgrad_mn4d(d) && d->k >= 2;
}
// N tile over Cout: 128 or 256 (each CTA holds whole 64-channel chunks of its half); the narrower tile runs the tensor pipe a little
// below the wide one, so it wins only when it saves padding
static int wgrad_cg2_bn(int cout) {
  return cout > 128 ? 256 : 128;           // a narrower LAST tile is chosen per tile inside the kernel (wg2_tile_bn)
}
static void wgrad_cg2_plan(const w2l_conv_desc* d, int* bn_out, int* pairs_out, int* rounds_out, int* zero_out) {
  const int bn = wgrad_cg2_bn(d->Cout);
  const int64_t tiles = (((int64_t)d->k * ((d->Cin + kBlockM - 1) / kBlockM) + 1) / 2) * ((d->Cout + bn - 1) / bn);
  const int64_t iters = (int64_t)d->B * ((d->T_out + kBlockK - 1) / kBlockK);
  int64_t pairs = gemm_sms() / 2;
  const int64_t min_iters = 32;
  if (tiles * iters / pairs < min_iters) pairs = tiles * iters / min_iters > 0 ? tiles * iters / min_iters : 1;
  const int64_t rounds = tiles / pairs;
  if (bn_out) *bn_out = bn;
  if (pairs_out) *pairs_out = (int)pairs;
  if (rounds_out) *rounds_out = (int)rounds;
  if (zero_out) *zero_out = (tiles - rounds * pairs) > 0;
}

static int launch_wgrad_cg2(const GemmParams& p, int pairs, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    W2L_CUDA(cudaFuncSetAttribute(conv_wgrad_cg2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kC2Smem));
    configured = true;
  }
  if (pairs < 1) return W2L_OK;
  conv_wgrad_cg2_kernel<<<2 * pairs, kGemmThreads, kC2Smem, st>>>(p);
  return after_launch("conv_wgrad_cg2_kernel");
}

}  // namespace w2l

extern "C" {

int w2l_conv1d_fwd(const void* x, const void* w, const float* bias, const float* scale, const float* shift, float* bn_stats,
                   void* y, const w2l_conv_desc* d, void* stream) {
  using namespace w2l;
  int rc = check_desc(d, "conv1d_fwd");
  if (rc) return rc;
  W2L_REQUIRE(x && w && y, "conv1d_fwd: null pointer");
  W2L_REQUIRE((scale == nullptr) == (shift == nullptr), "conv1d_fwd: scale and shift must be given together");
  W2L_REQUIRE(d->ldy >= d->Cout, "conv1d_fwd: ldy < Cout");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  {
    const char* e = getenv("W2L_DBG");
    p.dbg = e ? atoi(e) : 0;
  }
  set_operand_type(p, d);
  const uint64_t eb = p.tf32 ? 4 : 2;                   // bytes per operand element
  const bool cg2 = cg2_wanted() && !(p.dbg & 1) && !p.tf32;
  if (!cg2 && !p.tf32) plan_rings(p, d->k, d->dilation, true);
  {
    uint64_t dims[3] = {(uint64_t)d->Cin, (uint64_t)d->x_rows, (uint64_t)d->B};
    uint64_t str[2] = {(uint64_t)d->Cin * eb, (uint64_t)d->x_rows * d->Cin * eb};
    uint32_t box[3] = {(uint32_t)p.kblk, p.slab_rows ? (uint32_t)p.slab_rows : (uint32_t)kBlockM, 1};
    rc = make_tensor_map(&p.tmA, x, (int)eb, 3, dims, str, box, true);
    if (rc) return rc;
  }
  p.BN = cg2 ? cg2_bn(d->Cout_pad) : pick_bn(d->Cout_pad, 16);
  p.c2_npad = d->Cout_pad;
  {
    uint64_t dims[3] = {(uint64_t)d->Cin, (uint64_t)d->Cout_pad, (uint64_t)d->k};
    uint64_t str[2] = {(uint64_t)d->Cin * eb, (uint64_t)d->Cout_pad * d->Cin * eb};
    uint32_t box[3] = {(uint32_t)p.kblk, cg2 ? (uint32_t)p.BN / 2 : (p.dbg & 1) ? 64u : (uint32_t)p.BN, 1};
    rc = make_tensor_map(&p.tmB, w, (int)eb, 3, dims, str, box, true);
    if (rc) return rc;
  }
  p.B = d->B;
  p.m_tiles = (d->T_out + kBlockM - 1) / kBlockM;
  p.n_tiles = (d->Cout_pad + p.BN - 1) / p.BN;
  p.k = d->k;
  p.dil = d->dilation;
  p.kc_steps = (d->Cin + p.kblk - 1) / p.kblk;
  p.a_row_off = d->x_row_offset;
  p.a_tap_step = d->dilation;
  p.M_valid = d->T_out;
  p.N_valid = d->Cout;
  p.num_tiles = p.m_tiles * p.n_tiles * p.B;
  p.bias = bias;
  p.scale = scale;
  p.shift = shift;
  p.bn_stats = bn_stats;
  p.act = d->act;
  p.y_dtype = d->y_dtype;
  p.y = y;
  p.y_batch_stride = (int64_t)d->y_rows * d->ldy;
  p.y_row_off = d->y_row_offset;
  p.ldy = d->ldy;
  p.splits = 1;
  if (cg2) {
    plan_cg2(p, (int64_t)d->k * p.BN * d->Cin * 2);
    return launch_gemm_cg2(p, (cudaStream_t)stream);
  }
  plan_tail_split(p);
  return launch_gemm<MODE_FWD>(p, (cudaStream_t)stream);
}

int32_t w2l_conv1d_fwd_tail_parts(const w2l_conv_desc* d) {
  using namespace w2l;
  if (!d || d->B < 1 || d->T_out < 1 || d->k < 1 || d->Cin < 1 || d->Cout_pad < 16) return 0;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.BN = pick_bn(d->Cout_pad, 16);
  p.k = d->k;
  p.kc_steps = (d->Cin + kBlockK - 1) / kBlockK;
  p.num_tiles = ((d->T_out + kBlockM - 1) / kBlockM) * ((d->Cout_pad + p.BN - 1) / p.BN) * d->B;
  plan_tail_split(p);
  return p.tail_parts;
}

int w2l_set_gemm_scratch(void* scratch, size_t bytes) {
  using namespace w2l;
  W2L_REQUIRE((scratch == nullptr) == (bytes == 0), "set_gemm_scratch: pointer and size must be given together");
  W2L_REQUIRE(scratch == nullptr || (bytes > kScratchCounterBytes && ((uintptr_t)scratch & 255) == 0), "set_gemm_scratch: need > %zu bytes, 256-byte aligned",
              kScratchCounterBytes);
  g_scratch = scratch;
  g_scratch_bytes = bytes;
  return W2L_OK;
}

int w2l_conv1d_dgrad(const void* dy, const void* w, void* dx, const w2l_conv_desc* d, void* stream) {
  using namespace w2l;
  int rc = check_desc(d, "conv1d_dgrad");
  if (rc) return rc;
  W2L_REQUIRE(dy && w && dx, "conv1d_dgrad: null pointer");
  W2L_REQUIRE(d->x_dtype == W2L_DTYPE_BF16, "conv1d_dgrad: bf16 operands only (fp32 operands: w2l_conv1d_dgrad_wt)");
  W2L_REQUIRE(d->Cout_pad >= 64, "conv1d_dgrad: Cout_pad=%d must be >= 64 (pad dy/weights)", d->Cout_pad);
  W2L_REQUIRE(d->ldy >= d->Cout_pad, "conv1d_dgrad: dy row pitch %d < Cout_pad %d", d->ldy, d->Cout_pad);
  GemmParams p;
  memset(&p, 0, sizeof(p));
  {
    // dy viewed as [B, T_out, Cout_pad] inside a [B, y_rows, ldy] buffer
    const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(dy) + (int64_t)d->y_row_offset * d->ldy;
    uint64_t dims[3] = {(uint64_t)d->Cout_pad, (uint64_t)d->T_out, (uint64_t)d->B};
    uint64_t str[2] = {(uint64_t)d->ldy * 2, (uint64_t)d->y_rows * d->ldy * 2};
    uint32_t box[3] = {kBlockK, kBlockM, 1};
    rc = make_tensor_map(&p.tmA, base, 2, 3, dims, str, box, true);
    if (rc) return rc;
  }
  const int n_pad = (d->Cin + 15) / 16 * 16;
  p.BN = pick_bn(n_pad, 16);
  {
    uint64_t dims[3] = {(uint64_t)d->Cin, (uint64_t)d->Cout_pad, (uint64_t)d->k};
    uint64_t str[2] = {(uint64_t)d->Cin * 2, (uint64_t)d->Cout_pad * d->Cin * 2};
    uint32_t box[3] = {64, 64, 1};
    rc = make_tensor_map(&p.tmB, w, 2, 3, dims, str, box, true);
    if (rc) return rc;
  }
  p.B = d->B;
  p.m_tiles = (d->x_rows + kBlockM - 1) / kBlockM;
  p.n_tiles = (n_pad + p.BN - 1) / p.BN;
  p.k = d->k;
  p.dil = d->dilation;
  p.kc_steps = (d->Cout_pad + kBlockK - 1) / kBlockK;
  p.a_row_off = -d->x_row_offset;
  p.a_tap_step = -d->dilation;
  p.M_valid = d->x_rows;
  p.N_valid = d->Cin;
  p.num_tiles = p.m_tiles * p.n_tiles * p.B;
  p.act = W2L_ACT_NONE;
  p.y_dtype = W2L_DTYPE_BF16;
  p.y = dx;
  p.y_batch_stride = (int64_t)d->x_rows * d->Cin;
  p.y_row_off = 0;
  p.ldy = d->Cin;
  p.splits = 1;
  return launch_gemm<MODE_DGRAD>(p, (cudaStream_t)stream);
}

static int dgrad_wt_impl(const void* dy, const void* wt, void* dx, const w2l_conv_desc* d, const w2l_bn_reduce* red, void* stream) {
  // Backward-data as a FORWARD implicit GEMM over dy with tap-reversed, transposed weights wt[k-1-j][ci][co] = w[j][co][ci]:
  //   dx[u, ci] = sum_{j'} sum_co dy[u - off - (k-1)d + j'd, co] * wt[j'][ci][co]      (both operands K-major)
  using namespace w2l;
  int rc = check_desc(d, "conv1d_dgrad_wt");
  if (rc) return rc;
  W2L_REQUIRE(dy && wt && dx, "conv1d_dgrad_wt: null pointer");
  W2L_REQUIRE(d->Cout_pad >= 64 && d->Cout_pad % 8 == 0, "conv1d_dgrad_wt: Cout_pad=%d must be >= 64", d->Cout_pad);
  W2L_REQUIRE(d->ldy >= d->Cout, "conv1d_dgrad_wt: dy row pitch %d < Cout %d", d->ldy, d->Cout);
  GemmParams p;
  memset(&p, 0, sizeof(p));
  set_operand_type(p, d);
  const uint64_t eb = p.tf32 ? 4 : 2;
  const bool cg2 = cg2_wanted() && !p.tf32;
  if (!cg2 && !p.tf32) plan_rings(p, d->k, d->dilation, true);
  {
    // dy rows normally carry Cout_pad columns (zero padded); a row of only Cout columns (a hidden width that is a multiple of 8 but
    // not of 16) is declared as such, and the tail of the last contraction chunk reads as zero (TMA out-of-bounds fill)
    const char* base = reinterpret_cast<const char*>(dy) + (int64_t)d->y_row_offset * d->ldy * (int64_t)eb;
    const int a_cols = d->ldy >= d->Cout_pad ? d->Cout_pad : d->Cout;
    uint64_t dims[3] = {(uint64_t)a_cols, (uint64_t)d->T_out, (uint64_t)d->B};
    uint64_t str[2] = {(uint64_t)d->ldy * eb, (uint64_t)d->y_rows * d->ldy * eb};
    uint32_t box[3] = {(uint32_t)p.kblk, p.slab_rows ? (uint32_t)p.slab_rows : (uint32_t)kBlockM, 1};
    rc = make_tensor_map(&p.tmA, base, (int)eb, 3, dims, str, box, true);
    if (rc) return rc;
  }
  const int n_pad = (d->Cin + 15) / 16 * 16;
  p.BN = cg2 ? cg2_bn(n_pad) : pick_bn(n_pad, 16);
  p.c2_npad = n_pad;
  {
    uint64_t dims[3] = {(uint64_t)d->Cout_pad, (uint64_t)n_pad, (uint64_t)d->k};
    uint64_t str[2] = {(uint64_t)d->Cout_pad * eb, (uint64_t)n_pad * d->Cout_pad * eb};
    uint32_t box[3] = {(uint32_t)p.kblk, cg2 ? (uint32_t)p.BN / 2 : (uint32_t)p.BN, 1};
    rc = make_tensor_map(&p.tmB, wt, (int)eb, 3, dims, str, box, true);
    if (rc) return rc;
  }
  p.B = d->B;
  p.m_tiles = (d->x_rows + kBlockM - 1) / kBlockM;
  p.n_tiles = (n_pad + p.BN - 1) / p.BN;
  p.k = d->k;
  p.dil = d->dilation;
  p.kc_steps = (d->Cout_pad + p.kblk - 1) / p.kblk;
  p.a_row_off = -d->x_row_offset - (d->k - 1) * d->dilation;
  p.a_tap_step = d->dilation;
  p.M_valid = d->x_rows;
  p.N_valid = d->Cin;
  p.num_tiles = p.m_tiles * p.n_tiles * p.B;
  p.act = W2L_ACT_NONE;
  p.y_dtype = p.tf32 ? W2L_DTYPE_F32 : W2L_DTYPE_BF16;    // dx has the operand type
  p.y = dx;
  p.y_batch_stride = (int64_t)d->x_rows * d->Cin;
  p.y_row_off = 0;
  p.ldy = d->Cin;
  p.splits = 1;
  if (red != nullptr) {                // fold the BatchNorm-backward reduction of the block below into the epilogue
    const int Tp = red->pad_left + red->T + red->pad_right;
    W2L_REQUIRE(!p.tf32, "conv1d_dgrad_wt_bnred: bf16 activations only");
    W2L_REQUIRE(red->z && red->scale && red->shift && red->mean && red->red, "conv1d_dgrad_wt_bnred: null pointer");
    W2L_REQUIRE(red->B >= 1 && red->T >= 1 && red->pad_left >= 0 && red->pad_right >= 0 && red->pad_left < red->T && red->pad_right < red->T,
                "conv1d_dgrad_wt_bnred: bad block geometry");
    W2L_REQUIRE((int64_t)red->B * Tp == (int64_t)d->B * d->x_rows, "conv1d_dgrad_wt_bnred: %d x %d padded rows of the block do not match the %d x %d rows of dx",
                red->B, Tp, d->B, d->x_rows);
    W2L_REQUIRE(d->B == 1 || d->x_rows == Tp, "conv1d_dgrad_wt_bnred: dx rows per batch entry must be the block's padded rows (or one flat entry)");
    W2L_REQUIRE(red->act >= 0 && red->act <= 2 && red->drop_p >= 0.f && red->drop_p < 1.f, "conv1d_dgrad_wt_bnred: bad activation / dropout");
    W2L_REQUIRE(!(red->drop_p > 0.f) || red->drop_mask, "conv1d_dgrad_wt_bnred: dropout needs the stored keep-bits");
    p.r_z = reinterpret_cast<const __nv_bfloat16*>(red->z);
    p.r_mask = reinterpret_cast<const uint8_t*>(red->drop_mask);
    p.r_scale = red->scale;
    p.r_shift = red->shift;
    p.r_mean = red->mean;
    p.r_lens = red->lens;
    p.r_red = red->red;
    p.r_B = red->B;
    p.r_T = red->T;
    p.r_pl = red->pad_left;
    p.r_Tp = Tp;
    p.r_act = red->act;
    p.r_inv_keep = 1.f;
    if (red->drop_p > 0.f) {           // the passes' 12-bit keep probability (elementwise.cu drop_quant): the SAME quantisation
      const uint32_t one = 1u << 12;
      uint32_t q = (uint32_t)((1.0 - (double)red->drop_p) * (double)one + 0.5);
      if (q < 1) q = 1;
      if (q > one - 1) q = one - 1;
      p.r_inv_keep = (float)((double)one / (double)q);
    }
  }
  if (cg2) {
    plan_cg2(p, (int64_t)d->k * p.BN * d->Cout_pad * 2);
    return launch_gemm_cg2(p, (cudaStream_t)stream);
  }
  p.tail_first = p.num_tiles;          // backward-data overlaps with wgrad, whose CTAs fill its last wave: no tail split here
  return launch_gemm<MODE_FWD>(p, (cudaStream_t)stream);
}

int w2l_conv1d_dgrad_wt(const void* dy, const void* wt, void* dx, const w2l_conv_desc* d, void* stream) {
  return dgrad_wt_impl(dy, wt, dx, d, nullptr, stream);
}

int w2l_conv1d_dgrad_wt_bnred(const void* dy, const void* wt, void* dx, const w2l_conv_desc* d, const w2l_bn_reduce* red, void* stream) {
  W2L_REQUIRE(red != nullptr, "conv1d_dgrad_wt_bnred: null reduction descriptor");
  return dgrad_wt_impl(dy, wt, dx, d, red, stream);
}

int32_t w2l_conv1d_wgrad_splits(const w2l_conv_desc* d) {
  if (!d || d->B < 1 || d->T_out < 1 || d->k < 1) return 1;
  int zero = 0;
  if (w2l::wgrad_cg2_ok(d))
    w2l::wgrad_cg2_plan(d, nullptr, nullptr, nullptr, &zero);
  else
    w2l::wgrad_plan(d, nullptr, nullptr, nullptr, &zero);
  return zero ? 2 : 1;
}

int w2l_conv1d_wgrad_t(const float* dyT, int64_t dy_pitch, const float* const* xT_shifted, int64_t x_pitch, float* dw,
                       const w2l_conv_desc* d, void* stream) {
  // fp32-faithful weight gradient over TRANSPOSED fp32 operands (time contiguous), so that both are K-major like the forward
  // GEMM (an MN-major tf32 operand would need the 32-byte-atom swizzle):
  //   dw[j, co, ci] (+)= sum_b sum_t dyT[b, co, t] * xT[b, ci, t + x_row_offset + j*dilation]
  // xT_shifted[s] (HOST array of 4 device pointers) = xT delayed by s rows, xT_s[b, ci, u] = x[b, u - s, ci] with s zeros in front,
  // each with pitch x_pitch >= x_rows + s; only the residues s = (-(x_row_offset + j*dilation)) mod 4 that occur need to be non-null.
  using namespace w2l;
  int rc = check_desc(d, "conv1d_wgrad_t");
  if (rc) return rc;
  W2L_REQUIRE(dyT && xT_shifted && dw, "conv1d_wgrad_t: null pointer");
  for (int j = 0; j < d->k; ++j)
    W2L_REQUIRE(xT_shifted[(-(d->x_row_offset + j * d->dilation)) & 3] != nullptr, "conv1d_wgrad_t: tap %d needs the copy delayed by %d rows",
                j, (-(d->x_row_offset + j * d->dilation)) & 3);
  W2L_REQUIRE(d->x_dtype == W2L_DTYPE_F32, "conv1d_wgrad_t: fp32 operands only (x_dtype)");
  W2L_REQUIRE(dy_pitch >= d->T_out && x_pitch >= d->x_rows + 3 && dy_pitch % 4 == 0 && x_pitch % 4 == 0,
              "conv1d_wgrad_t: pitches must cover the rows and be multiples of 4 floats (dy_pitch %lld, x_pitch %lld)",
              (long long)dy_pitch, (long long)x_pitch);
  GemmParams p;
  memset(&p, 0, sizeof(p));
  set_operand_type(p, d);
  p.wg_kmajor = 1;
  const int n_pad = (d->Cin + 15) / 16 * 16;
  p.BN = pick_bn(n_pad, 16);
  {
    uint64_t dims[3] = {(uint64_t)d->T_out, (uint64_t)d->Cout, (uint64_t)d->B};
    uint64_t str[2] = {(uint64_t)dy_pitch * 4, (uint64_t)d->Cout * dy_pitch * 4};
    uint32_t box[3] = {(uint32_t)p.kblk, kBlockM, 1};
    rc = make_tensor_map(&p.tmA, dyT, 4, 3, dims, str, box, true);
    if (rc) return rc;
  }
  p.tmB = p.tmA;                                                   // (prefetched by the kernel; never loaded through in this mode)
  for (int sft = 0; sft < 4; ++sft) {
    if (!xT_shifted[sft]) {
      p.tmX[sft] = p.tmA;
      continue;
    }
    uint64_t dims[3] = {(uint64_t)(d->x_rows + sft), (uint64_t)d->Cin, (uint64_t)d->B};
    uint64_t str[2] = {(uint64_t)x_pitch * 4, (uint64_t)d->Cin * x_pitch * 4};
    uint32_t box[3] = {(uint32_t)p.kblk, (uint32_t)p.BN, 1};
    rc = make_tensor_map(&p.tmX[sft], xT_shifted[sft], 4, 3, dims, str, box, true);
    if (rc) return rc;
  }
  p.B = d->B;
  p.m_tiles = (d->Cout + kBlockM - 1) / kBlockM;
  p.n_tiles = (n_pad + p.BN - 1) / p.BN;
  p.k = d->k;
  p.dil = d->dilation;
  p.kc_steps = (d->T_out + p.kblk - 1) / p.kblk;
  p.b_row_off = d->x_row_offset;
  p.M_valid = d->Cout;
  p.N_valid = d->Cin;
  p.num_tiles = p.k * p.m_tiles * p.n_tiles;
  p.y = dw;
  p.ldy = d->Cin;
  p.dw_tap_stride = (int64_t)d->Cout * d->Cin;
  // stream-K plan over the single-CTA grid; dw must be ZERO on entry (tiles shared between CTAs are accumulated with atomics)
  const int64_t iters = (int64_t)p.B * p.kc_steps;
  int64_t grid = gemm_sms();
  if ((int64_t)p.num_tiles * iters / grid < 32) grid = (int64_t)p.num_tiles * iters / 32 > 0 ? (int64_t)p.num_tiles * iters / 32 : 1;
  p.wg_rounds = (int32_t)(p.num_tiles / grid);
  p.splits = 2;
  return launch_gemm<MODE_WGRAD>(p, (cudaStream_t)stream, (int)grid);
}

int w2l_conv1d_wgrad(const void* dy, const void* x, float* dw, const w2l_conv_desc* d, void* stream) {
  using namespace w2l;
  int rc = check_desc(d, "conv1d_wgrad");
  if (rc) return rc;
  W2L_REQUIRE(dy && x && dw, "conv1d_wgrad: null pointer");
  W2L_REQUIRE(d->x_dtype == W2L_DTYPE_BF16, "conv1d_wgrad: bf16 operands only (fp32 operands: w2l_conv1d_wgrad_t)");
  W2L_REQUIRE(d->ldy >= 64 && d->ldy % 8 == 0, "conv1d_wgrad: dy row pitch %d must be >= 64 and a multiple of 8", d->ldy);
  GemmParams p;
  memset(&p, 0, sizeof(p));
  const __nv_bfloat16* dy_base = reinterpret_cast<const __nv_bfloat16*>(dy) + (int64_t)d->y_row_offset * d->ldy;
  if (wgrad_cg2_ok(d)) {
    // CTA-pair kernel on the transposed problem: A = x (128 ci x 64 rows, shifted per tap), B = dy (BN co x 64 rows, half per CTA)
    int bn = 0, pairs = 0, rounds = 0, zero = 0;
    wgrad_cg2_plan(d, &bn, &pairs, &rounds, &zero);
    {
      uint64_t dims[4] = {64, (uint64_t)d->x_rows, (uint64_t)d->Cin / 64, (uint64_t)d->B};
      uint64_t str[3] = {(uint64_t)d->Cin * 2, 128, (uint64_t)d->x_rows * d->Cin * 2};
      uint32_t box[4] = {64, kBlockK, 2, 1};
      rc = make_tensor_map(&p.tmA, x, 2, 4, dims, str, box, true);
      if (rc) return rc;
    }
    {
      uint64_t dims[4] = {64, (uint64_t)d->T_out, (uint64_t)d->ldy / 64, (uint64_t)d->B};
      uint64_t str[3] = {(uint64_t)d->ldy * 2, 128, (uint64_t)d->y_rows * d->ldy * 2};
      uint32_t box[4] = {64, kBlockK, (uint32_t)bn / 128, 1};
      rc = make_tensor_map(&p.tmB, dy_base, 2, 4, dims, str, box, true);
      if (rc) return rc;
    }
    p.BN = bn;
    p.splits = zero ? 2 : 1;
    p.wg_rounds = rounds;
    p.B = d->B;
    p.m_tiles = 1;                        // the (ci tile, tap) axis is flattened: UnitIter decodes p.k PAIRS of it per co tile
    p.n_tiles = (d->Cout + bn - 1) / bn;
    p.k = (d->k * ((d->Cin + kBlockM - 1) / kBlockM) + 1) / 2;
    p.c2_k = d->k;
    p.dil = d->dilation;
    p.kc_steps = (d->T_out + kBlockK - 1) / kBlockK;
    p.b_row_off = d->x_row_offset;
    p.M_valid = d->Cin;
    p.N_valid = d->Cout;
    p.num_tiles = p.k * p.m_tiles * p.n_tiles;
    p.y = dw;
    p.ldy = d->Cin;
    p.dw_tap_stride = (int64_t)d->Cout * d->Cin;
    return launch_wgrad_cg2(p, pairs, (cudaStream_t)stream);
  }
  p.mn4d = wgrad_mn4d(d);
  int bn = 0, grid = 0, rounds = 0, zero = 0;
  wgrad_plan(d, &bn, &grid, &rounds, &zero);
  if (p.mn4d) {
    {
      uint64_t dims[4] = {64, (uint64_t)d->T_out, (uint64_t)d->ldy / 64, (uint64_t)d->B};
      uint64_t str[3] = {(uint64_t)d->ldy * 2, 128, (uint64_t)d->y_rows * d->ldy * 2};
      uint32_t box[4] = {64, kBlockK, 2, 1};
      rc = make_tensor_map(&p.tmA, dy_base, 2, 4, dims, str, box, true);
      if (rc) return rc;
    }
    {
      uint64_t dims[4] = {64, (uint64_t)d->x_rows, (uint64_t)d->Cin / 64, (uint64_t)d->B};
      uint64_t str[3] = {(uint64_t)d->Cin * 2, 128, (uint64_t)d->x_rows * d->Cin * 2};
      uint32_t box[4] = {64, kBlockK, (uint32_t)((bn + 63) / 64), 1};
      rc = make_tensor_map(&p.tmB, x, 2, 4, dims, str, box, true);
      if (rc) return rc;
    }
  } else {
    {
      uint64_t dims[3] = {(uint64_t)d->Cout, (uint64_t)d->T_out, (uint64_t)d->B};   // columns >= Cout read as zero
      uint64_t str[2] = {(uint64_t)d->ldy * 2, (uint64_t)d->y_rows * d->ldy * 2};
      uint32_t box[3] = {64, kBlockK, 1};
      rc = make_tensor_map(&p.tmA, dy_base, 2, 3, dims, str, box, true);
      if (rc) return rc;
    }
    {
      uint64_t dims[3] = {(uint64_t)d->Cin, (uint64_t)d->x_rows, (uint64_t)d->B};
      uint64_t str[2] = {(uint64_t)d->Cin * 2, (uint64_t)d->x_rows * d->Cin * 2};
      uint32_t box[3] = {64, kBlockK, 1};
      rc = make_tensor_map(&p.tmB, x, 2, 3, dims, str, box, true);
      if (rc) return rc;
    }
  }
  p.BN = bn;
  p.splits = zero ? 2 : 1;
  p.wg_rounds = rounds;
  const int n_pad = (d->Cin + 15) / 16 * 16;
  p.B = d->B;
  p.m_tiles = (d->Cout + kBlockM - 1) / kBlockM;
  p.n_tiles = (n_pad + p.BN - 1) / p.BN;
  p.k = d->k;
  p.dil = d->dilation;
  p.kc_steps = (d->T_out + kBlockK - 1) / kBlockK;
  p.b_row_off = d->x_row_offset;
  p.M_valid = d->Cout;
  p.N_valid = d->Cin;
  p.num_tiles = p.k * p.m_tiles * p.n_tiles;
  p.y = dw;
  p.ldy = d->Cin;
  p.dw_tap_stride = (int64_t)d->Cout * d->Cin;
  return launch_gemm<MODE_WGRAD>(p, (cudaStream_t)stream, grid);
}

}  // extern "C"
