// Hardware probe for the round-2 "resident activation slab" GEMM: can tcgen05.mma read a K-major,
// 128B-swizzled A operand that starts at an ARBITRARY row of a larger shared-memory slab (start address = slab + r0 * 128 B,
// i.e. not aligned to the 1024-byte swizzle atom), and does the descriptor's base_offset field (bits 49-51) have to carry
// r0 & 7 for that?  One CTA: slab of 192 rows x 64 bf16, B tile 64 x 64, D[128 x 64] = A[r0 : r0+128] . B^T for r0 = 0..15,
// once with base_offset = 0 and once with base_offset = r0 & 7, compared with a host reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/probe_umma_row_offset tools/probe_umma_row_offset.cu && ./tools/probe_umma_row_offset
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "../include/w2l_sm100.h"
#include "../wav2letter_pytorch_b200/csrc/common.cuh"

using namespace w2l;

constexpr int kSlabRows = 192, kK = 64, kN = 64, kM = 128;

__device__ __forceinline__ uint64_t desc_with_base(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((16u >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_offset & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(128) probe_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                                    float* __restrict__ d_out, int r0, int use_base) {
  extern __shared__ uint8_t raw[];
  const uint32_t raw_addr = smem_u32(raw);
  uint8_t* smem = raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sa = smem;                                  // slab: 192 rows x 128 B
  uint8_t* sb = smem + kSlabRows * 128;                // 64 rows x 128 B (24576 is a multiple of 1024)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sb + kN * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // K-major, SWIZZLE_128B: 16-byte chunk c16 of row r lives at r*128 + ((c16 ^ (r & 7)) * 16)
  for (int i = tid; i < kSlabRows * 8; i += 128) {
    const int r = i >> 3, c16 = i & 7;
    *reinterpret_cast<uint4*>(sa + r * 128 + ((c16 ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(a + r * kK + c16 * 8);
  }
  for (int i = tid; i < kN * 8; i += 128) {
    const int r = i >> 3, c16 = i & 7;
    *reinterpret_cast<uint4*>(sb + r * 128 + ((c16 ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(b + r * kK + c16 * 8);
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 64);
    tmem_relinquish();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core's async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(kM, kN, false, false);
    const uint32_t a_addr = smem_u32(sa) + (uint32_t)r0 * 128u, b_addr = smem_u32(sb);
    for (int kk = 0; kk < kK / 16; ++kk) {
      const uint64_t adesc = desc_with_base(a_addr + kk * 32, 1024u, use_base ? (uint32_t)(r0 & 7) : 0u);
      const uint64_t bdesc = desc_with_base(b_addr + kk * 32, 1024u, 0u);
      umma_bf16(tmem_base, adesc, bdesc, idesc, kk > 0);
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < kN; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) d_out[(warp * 32 + lane) * kN + c0 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

int main() {
  std::vector<__nv_bfloat16> ha(kSlabRows * kK), hb(kN * kK);
  std::vector<float> fa(kSlabRows * kK), fb(kN * kK);
  uint32_t s = 12345u;
  auto rnd = [&]() {
    s = s * 1664525u + 1013904223u;
    return (float)((int)((s >> 9) & 0xFF) - 128) / 64.f;       // small multiples of 1/64: exact in bf16, exact fp32 sums
  };
  for (size_t i = 0; i < ha.size(); ++i) {
    fa[i] = rnd();
    ha[i] = __float2bfloat16(fa[i]);
  }
  for (size_t i = 0; i < hb.size(); ++i) {
    fb[i] = rnd();
    hb[i] = __float2bfloat16(fb[i]);
  }
  __nv_bfloat16 *da, *db;
  float* dd;
  cudaMalloc(&da, ha.size() * 2);
  cudaMalloc(&db, hb.size() * 2);
  cudaMalloc(&dd, kM * kN * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  const size_t smem = kSlabRows * 128 + kN * 128 + 64 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  std::vector<float> hd(kM * kN);
  printf("r0  base_offset=0: max|err|   base_offset=r0&7: max|err|\n");
  for (int r0 = 0; r0 < 16; ++r0) {
    double err[2];
    for (int ub = 0; ub < 2; ++ub) {
      cudaMemset(dd, 0, kM * kN * 4);
      probe_kernel<<<1, 128, smem>>>(da, db, dd, r0, ub);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("r0=%d use_base=%d: CUDA error %s\n", r0, ub, cudaGetErrorString(e));
        return 1;
      }
      cudaMemcpy(hd.data(), dd, kM * kN * 4, cudaMemcpyDeviceToHost);
      double worst = 0;
      for (int m = 0; m < kM; ++m)
        for (int n = 0; n < kN; ++n) {
          double ref = 0;
          for (int k = 0; k < kK; ++k) ref += (double)fa[(r0 + m) * kK + k] * (double)fb[n * kK + k];
          worst = fmax(worst, fabs(ref - (double)hd[m * kN + n]));
        }
      err[ub] = worst;
    }
    printf("%2d  %22.6f   %24.6f\n", r0, err[0], err[1]);
  }
  return 0;
}
