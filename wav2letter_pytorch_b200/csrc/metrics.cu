// GPU-side WER/CER bookkeeping for the training step.
// The reference decodes and scores every step on the host (base_asr_models.py:53-69: GreedyDecoder.decode, then
// Levenshtein per utterance), which on a GPU means a device sync plus milliseconds of idle accelerator per step.  Here the
// collapsed token ids never leave the device: they are split into the character stream without spaces (CER) and a stream of
// 64-bit word hashes (WER), scored against the encoded reference by a wavefront Levenshtein kernel, and reduced to the three
// logged ratios.  The host only reads the result when (if) it formats the log.
#include "common.cuh"

namespace w2l {

__device__ __forceinline__ unsigned long long hash_step(unsigned long long h, int sym) {
  return (h ^ (unsigned long long)(unsigned)(sym + 1)) * 1099511628211ull;     // FNV-1a over symbol ids
}
constexpr unsigned long long kHashSeed = 1469598103934665603ull;

// One CTA per utterance: tokens[n, :counts[n]] -> chars (ids != space, order kept) and word hashes (one per run of ids != space,
// order kept).  Stream compaction over 128-token chunks: warp ballots give every symbol its rank among the chunk's characters
// and every word start its rank among the word starts; the thread that owns a word start walks its word (words are short) and
// writes the hash.  (Round 1 ran one THREAD per utterance: a 750-step dependent loop, 129 us at N=64 x T=750.)
constexpr int kSplitThreads = 128;
__global__ void __launch_bounds__(kSplitThreads)
metrics_split_kernel(const int32_t* __restrict__ tokens, const int32_t* __restrict__ counts, int N, int T, int space,
                     int32_t* __restrict__ chars, int32_t* __restrict__ n_chars, long long* __restrict__ words,
                     int32_t* __restrict__ n_words) {
  __shared__ int s_c[kSplitThreads / 32], s_w[kSplitThreads / 32];
  const int n = blockIdx.x;
  if (n >= N) return;
  const int32_t* tk = tokens + (int64_t)n * T;
  int32_t* ch = chars + (int64_t)n * T;
  long long* wd = words + (int64_t)n * T;
  const int cnt = max(0, min(T, counts[n]));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned below = (1u << lane) - 1u;
  int nc = 0, nw = 0;                                    // running totals, the same in every thread
  for (int base = 0; base < cnt; base += kSplitThreads) {
    const int i = base + (int)threadIdx.x;
    const int s = i < cnt ? tk[i] : space;
    const bool is_char = i < cnt && s != space;
    const bool is_start = is_char && (i == 0 || tk[i - 1] == space);
    const unsigned mc = __ballot_sync(0xffffffffu, is_char), mw = __ballot_sync(0xffffffffu, is_start);
    if (lane == 0) {
      s_c[warp] = __popc(mc);
      s_w[warp] = __popc(mw);
    }
    __syncthreads();
    int oc = nc + __popc(mc & below), ow = nw + __popc(mw & below);
#pragma unroll
    for (int w = 0; w < kSplitThreads / 32; ++w) {
      if (w < warp) {
        oc += s_c[w];
        ow += s_w[w];
      }
      nc += s_c[w];
      nw += s_w[w];
    }
    if (is_char) ch[oc] = s;
    if (is_start) {
      unsigned long long h = kHashSeed;
      for (int k = i; k < cnt; ++k) {
        const int v = tk[k];
        if (v == space) break;
        h = hash_step(h, v);
      }
      wd[ow] = (long long)h;
    }
    __syncthreads();                                     // s_c / s_w are rewritten by the next chunk
  }
  if (threadIdx.x == 0) {
    n_chars[n] = nc;
    n_words[n] = nw;
  }
}

// Levenshtein distance, one WARP per (hypothesis, reference) pair: lane l owns the K reference columns l*K+1 .. l*K+K of the DP table
// in registers and sweeps the hypothesis rows one step behind lane l-1 (at step s lane l is on row s-l+1), so the only traffic between
// lanes is one shuffle per step -- the last column of the left neighbour's strip.  m + (n-1)/K steps, each a log2(K)-deep prefix
// minimum over the strip; no shared memory, no block barriers.  (Rounds 1-2 ran one CTA per pair, a __syncthreads per anti-diagonal: 90 us for 225-symbol
// pairs, and barrier-heavy CTAs beside the CTC recursion.)  dist(a[0:m], b[0:n]); n <= 32*K.
constexpr int kEditWarps = 4;
template <typename Sym, int K>
__global__ void __launch_bounds__(32 * kEditWarps)
edit_distance_kernel(const Sym* __restrict__ hyp, const int32_t* __restrict__ hyp_len, int64_t hyp_stride, const Sym* __restrict__ ref,
                     const int32_t* __restrict__ ref_len, int64_t ref_stride, int32_t* __restrict__ out, int pairs) {
  const int pair = blockIdx.x * kEditWarps + (int)(threadIdx.x >> 5);
  if (pair >= pairs) return;                         // (the whole warp leaves together)
  const int lane = threadIdx.x & 31;
  const int m = hyp_len[pair], n = ref_len[pair];
  const Sym* a = hyp + (int64_t)pair * hyp_stride;
  const Sym* b = ref + (int64_t)pair * ref_stride;
  if (m == 0 || n == 0) {
    if (lane == 0) out[pair] = m + n;
    return;
  }
  Sym bq[K];
  int row[K];                                        // D[i-1][c] of the strip while the lane works on row i; D[0][c] = c to start with
#pragma unroll
  for (int q = 0; q < K; ++q) {
    const int c = lane * K + q + 1;
    bq[q] = c <= n ? b[c - 1] : Sym(0);
    row[q] = c;
  }
  const int owner = (n - 1) / K;                     // the lane that holds column n
  int left_prev = lane * K;                          // D[i-1][lane*K], the column left of the strip: D[0][lane*K] for row 1
  int my_last = 0;                                   // D[i][lane*K + K] of the row finished in the previous step
  Sym a_next = lane == 0 ? a[0] : Sym(0);            // hypothesis symbol of the row this lane starts next, fetched a step ahead
  for (int s = 0; s < m + owner; ++s) {
    const int recv = __shfl_up_sync(0xffffffffu, my_last, 1);
    const int i = s - lane + 1;                      // 1-based row of this lane at this step
    const Sym ai = a_next;
    a_next = a[min(max(i, 0), m - 1)];               // row i+1's symbol (clamped: an unconditional load keeps the address math out of the loop)
    if (i >= 1 && i <= m) {
      // v_q = min(t_q, v_{q-1} + 1) with t_q = min(D[i-1][c] + 1, D[i-1][c-1] + cost) and v_{-1} = D[i][lane*K] (from the left lane).
      // Unrolled: v_q = q + min(left + 1, min_{j<=q}(t_j - j)) -- the t_j and their prefix minimum depend on the lane's own previous
      // row only, so what waits for the neighbour's shuffle is one add, one min, one add instead of a K-cell chain.
      const int left = lane == 0 ? i : recv;         // D[i][lane*K]
      int pm[K];
#pragma unroll
      for (int q = 0; q < K; ++q) {
        const int diag = q == 0 ? left_prev : row[q - 1];            // D[i-1][c-1]
        pm[q] = min(row[q] + 1, diag + (ai != bq[q] ? 1 : 0)) - q;
      }
      left_prev = left;
#pragma unroll
      for (int d = 1; d < K; d <<= 1) {              // inclusive prefix minimum (Kogge-Stone, in place from the top)
#pragma unroll
        for (int q = K - 1; q >= d; --q) pm[q] = min(pm[q], pm[q - d]);
      }
      const int l1 = left + 1;
#pragma unroll
      for (int q = 0; q < K; ++q) row[q] = min(l1, pm[q]) + q;
      my_last = row[K - 1];
    }
  }
  int res = 0;
#pragma unroll
  for (int q = 0; q < K; ++q)
    if (q == (n - 1) % K) res = row[q];
  res = __shfl_sync(0xffffffffu, res, owner);
  if (lane == 0) out[pair] = res;
}

// ratios[0] = cer, [1] = wer, [2] = len_ratio
__global__ void metrics_finalize_kernel(const int32_t* __restrict__ cer_d, const int32_t* __restrict__ wer_d,
                                        const int32_t* __restrict__ counts, int N, float cer_den, float wer_den, float len_den,
                                        float* __restrict__ ratios) {
  __shared__ long long s[3][32];
  long long c = 0, w = 0, l = 0;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    c += cer_d[n];
    w += wer_d[n];
    l += counts[n];
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    c += __shfl_xor_sync(0xffffffffu, c, o);
    w += __shfl_xor_sync(0xffffffffu, w, o);
    l += __shfl_xor_sync(0xffffffffu, l, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s[0][threadIdx.x >> 5] = c;
    s[1][threadIdx.x >> 5] = w;
    s[2][threadIdx.x >> 5] = l;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long tc = 0, tw = 0, tl = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
      tc += s[0][i];
      tw += s[1][i];
      tl += s[2][i];
    }
    ratios[0] = (float)((double)tc / (double)cer_den);
    ratios[1] = (float)((double)tw / (double)wer_den);
    ratios[2] = (float)((double)tl / (double)len_den);
  }
}

}  // namespace w2l

extern "C" {

size_t w2l_string_metrics_workspace_bytes(int64_t N, int64_t T, int64_t ref_stride) {
  // words (hyp [N,T] + ref [N,S]) i64, chars (hyp + ref) i32, 6 x [N] i32 counters
  return (size_t)((N * T + N * ref_stride) * 12 + 6 * N * 4 + 256);
}

int w2l_string_metrics(const int32_t* tokens, const int32_t* counts, int64_t N, int64_t T, int32_t space_index,
                       const int32_t* ref_ids, const int32_t* ref_lens, int64_t ref_stride, float cer_den, float wer_den,
                       float len_den, float* ratios, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace w2l;
  W2L_REQUIRE(tokens && counts && ref_ids && ref_lens && ratios && workspace, "string_metrics: null pointer");
  W2L_REQUIRE(N >= 1 && T >= 1, "string_metrics: bad shape");
  W2L_REQUIRE(ref_stride >= 1 && ref_stride <= 1023, "string_metrics: references longer than 1023 symbols are not supported on the device path");
  W2L_REQUIRE(workspace_bytes >= w2l_string_metrics_workspace_bytes(N, T, ref_stride), "string_metrics: workspace too small");
  W2L_REQUIRE(((uintptr_t)workspace & 7) == 0, "string_metrics: workspace must be 8-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  long long* h_words = (long long*)ws;
  long long* r_words = h_words + N * T;
  int32_t* h_chars = (int32_t*)(r_words + N * ref_stride);
  int32_t* r_chars = h_chars + N * T;
  int32_t* h_nc = r_chars + N * ref_stride;
  int32_t* h_nw = h_nc + N;
  int32_t* r_nc = h_nw + N;
  int32_t* r_nw = r_nc + N;
  int32_t* cer_d = r_nw + N;
  int32_t* wer_d = cer_d + N;
  metrics_split_kernel<<<(unsigned)N, kSplitThreads, 0, st>>>(tokens, counts, (int)N, (int)T, space_index, h_chars, h_nc, h_words, h_nw);
  int rc = after_launch("metrics_split_kernel<hyp>");
  if (rc) return rc;
  metrics_split_kernel<<<(unsigned)N, kSplitThreads, 0, st>>>(ref_ids, ref_lens, (int)N, (int)ref_stride, space_index, r_chars, r_nc, r_words, r_nw);
  rc = after_launch("metrics_split_kernel<ref>");
  if (rc) return rc;
  const unsigned eblocks = (unsigned)((N + kEditWarps - 1) / kEditWarps);
  if (ref_stride <= 256) {                          // 8 columns per lane
    edit_distance_kernel<int32_t, 8><<<eblocks, 32 * kEditWarps, 0, st>>>(h_chars, h_nc, T, r_chars, r_nc, ref_stride, cer_d, (int)N);
    rc = after_launch("edit_distance_kernel<chars>");
    if (rc) return rc;
    edit_distance_kernel<long long, 8><<<eblocks, 32 * kEditWarps, 0, st>>>(h_words, h_nw, T, r_words, r_nw, ref_stride, wer_d, (int)N);
  } else {                                          // up to 1023 reference symbols: 32 columns per lane
    edit_distance_kernel<int32_t, 32><<<eblocks, 32 * kEditWarps, 0, st>>>(h_chars, h_nc, T, r_chars, r_nc, ref_stride, cer_d, (int)N);
    rc = after_launch("edit_distance_kernel<chars>");
    if (rc) return rc;
    edit_distance_kernel<long long, 32><<<eblocks, 32 * kEditWarps, 0, st>>>(h_words, h_nw, T, r_words, r_nw, ref_stride, wer_d, (int)N);
  }
  rc = after_launch("edit_distance_kernel<words>");
  if (rc) return rc;
  metrics_finalize_kernel<<<1, 256, 0, st>>>(cer_d, wer_d, counts, (int)N, cer_den, wer_den, len_den, ratios);
  return after_launch("metrics_finalize_kernel");
}

}  // extern "C"
