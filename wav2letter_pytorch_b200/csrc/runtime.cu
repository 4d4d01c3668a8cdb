// Host-side runtime of libw2l_sm100: error text, launch counter, TMA descriptor encoding.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace w2l {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return W2L_OK;
  set_error("CUDA error %s (%d) at %s", cudaGetErrorString(e), (int)e, what);
  return W2L_ERR_CUDA;
}

int after_launch(const char* kernel_name) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("launch of %s failed: %s (%d)", kernel_name, cudaGetErrorString(e), (int)e);
    return W2L_ERR_CUDA;
  }
  return W2L_OK;
}

int num_sms() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    cached = n;
  }
  return cached;
}

// SMs the persistent GEMM kernels may occupy (w2l_set_sm_budget): the data-parallel host reserves a few SMs for the
// gradient collective, whose CTAs cannot co-reside with a 197 KB-shared-memory GEMM CTA -- without the reservation a
// 148-CTA persistent grid runs as two waves whenever a collective holds an SM.
static std::atomic<int> g_sm_budget{0};
int gemm_sms() {
  const int n = num_sms(), b = g_sm_budget.load(std::memory_order_relaxed);
  return (b > 0 && b < n) ? b : n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

int make_tensor_map(CUtensorMap* map, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available (no CUDA driver?)");
    return W2L_ERR_CUDA;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = enc(map, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu] box [%u,%u,%u] stride1 %llu base %p", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), (unsigned long long)(rank > 2 ? dims[2] : 0),
              box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), base);
    return W2L_ERR_INVALID_ARGUMENT;
  }
  return W2L_OK;
}

}  // namespace w2l

extern "C" {
int w2l_version(void) { return 100; }
#ifndef W2L_SRC_HASH
#define W2L_SRC_HASH "unknown"
#endif
// sha256 prefix of csrc/* + include/w2l_sm100.h at build time (_lib.source_hash): the loader refuses a library whose
// sources are not the ones in the tree, so a measured number always belongs to the kernels of the commit that reports it
static const char kSrcHash[] = "W2L_SRC_HASH=" W2L_SRC_HASH;
const char* w2l_source_hash(void) { return kSrcHash + 13; }
const char* w2l_last_error(void) { return w2l::g_err; }
int64_t w2l_launch_count(void) { return w2l::g_launches.load(); }
}

// ---------------------------------------------------------------- host-side Levenshtein (WER/CER bookkeeping)
// The reference scores every training step's transcripts with the python-Levenshtein C extension
// (decoder.py:31-66, base_asr_models.py:58-69); this is the same dynamic programme over int32 symbol ids.
extern "C" int64_t w2l_edit_distance_host(const int32_t* a, int64_t n, const int32_t* b, int64_t m) {
  if (n < 0 || m < 0 || (n && !a) || (m && !b)) return -1;
  if (n < m) {
    const int32_t* t = a; a = b; b = t;
    int64_t q = n; n = m; m = q;
  }
  if (m == 0) return n;
  int64_t stack_row[1024];
  int64_t* row = m + 1 <= 1024 ? stack_row : new int64_t[m + 1];
  for (int64_t j = 0; j <= m; ++j) row[j] = j;
  for (int64_t i = 1; i <= n; ++i) {
    int64_t diag = row[0];
    row[0] = i;
    const int32_t ai = a[i - 1];
    for (int64_t j = 1; j <= m; ++j) {
      const int64_t up = row[j];
      int64_t best = diag + (ai != b[j - 1]);
      if (up + 1 < best) best = up + 1;
      if (row[j - 1] + 1 < best) best = row[j - 1] + 1;
      row[j] = best;
      diag = up;
    }
  }
  const int64_t r = row[m];
  if (row != stack_row) delete[] row;
  return r;
}

// ---------------------------------------------------------------- batched, bit-parallel Levenshtein
// Myers' bit-vector algorithm in Hyyro's block formulation: O(ceil(m/64) * n) word operations per pair instead of
// O(m*n) cells; pairs are spread over host threads.  One call scores a whole batch of transcripts.
#include <thread>
#include <unordered_map>
#include <vector>

namespace w2l {
static int64_t myers_distance(const int32_t* a, int64_t m, const int32_t* b, int64_t n) {
  if (m == 0) return n;
  if (n == 0) return m;
  const int64_t words = (m + 63) / 64;
  std::unordered_map<int32_t, int32_t> ids;
  ids.reserve((size_t)m * 2);
  for (int64_t i = 0; i < m; ++i) ids.emplace(a[i], (int32_t)ids.size());
  std::vector<uint64_t> peq(ids.size() * (size_t)words, 0);
  for (int64_t i = 0; i < m; ++i) peq[(size_t)ids[a[i]] * words + i / 64] |= 1ull << (i % 64);
  std::vector<uint64_t> vp((size_t)words, ~0ull), vn((size_t)words, 0ull);
  const std::vector<uint64_t> zero((size_t)words, 0ull);
  const uint64_t last = 1ull << ((m - 1) % 64);
  int64_t dist = m;
  for (int64_t j = 0; j < n; ++j) {
    auto it = ids.find(b[j]);
    const uint64_t* pm = it == ids.end() ? zero.data() : &peq[(size_t)it->second * words];
    uint64_t hp_carry = 1, hn_carry = 0;
    for (int64_t w = 0; w < words; ++w) {
      const uint64_t x = pm[w] | hn_carry;
      const uint64_t d0 = (((x & vp[w]) + vp[w]) ^ vp[w]) | x | vn[w];
      uint64_t hp = vn[w] | ~(d0 | vp[w]);
      uint64_t hn = d0 & vp[w];
      if (w == words - 1) {
        dist += (hp & last) != 0;
        dist -= (hn & last) != 0;
      }
      const uint64_t hp_out = hp >> 63, hn_out = hn >> 63;
      hp = (hp << 1) | hp_carry;
      hn = (hn << 1) | hn_carry;
      hp_carry = hp_out;
      hn_carry = hn_out;
      vp[w] = hn | ~(d0 | hp);
      vn[w] = hp & d0;
    }
  }
  return dist;
}
}  // namespace w2l

extern "C" int w2l_edit_distance_batch_host(const int32_t* a, const int64_t* a_off, const int32_t* b, const int64_t* b_off,
                                            int64_t n_pairs, int64_t* out, int32_t threads) {
  if (n_pairs < 0 || !a_off || !b_off || !out) return W2L_ERR_INVALID_ARGUMENT;
  if (n_pairs == 0) return W2L_OK;
  auto work = [&](int64_t lo, int64_t hi) {
    for (int64_t p = lo; p < hi; ++p)
      out[p] = w2l::myers_distance(a + a_off[p], a_off[p + 1] - a_off[p], b + b_off[p], b_off[p + 1] - b_off[p]);
  };
  int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  if (nt > 16) nt = 16;
  if (nt > n_pairs) nt = (int)n_pairs;
  if (nt <= 1) {
    work(0, n_pairs);
    return W2L_OK;
  }
  std::vector<std::thread> pool;
  const int64_t per = (n_pairs + nt - 1) / nt;
  for (int t = 0; t < nt; ++t) {
    const int64_t lo = t * per, hi = lo + per < n_pairs ? lo + per : n_pairs;
    if (lo < hi) pool.emplace_back(work, lo, hi);
  }
  for (auto& th : pool) th.join();
  return W2L_OK;
}

extern "C" int w2l_set_sm_budget(int32_t sms) {
  w2l::g_sm_budget.store(sms > 0 ? sms : 0);
  return W2L_OK;
}
extern "C" int32_t w2l_get_sm_budget(void) { return w2l::gemm_sms(); }
