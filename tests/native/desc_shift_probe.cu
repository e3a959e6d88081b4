// Probe (run on a B200): does a K-major swizzled tcgen05 shared-memory descriptor whose start address is advanced by
// whole ROWS (s * row_bytes, not a multiple of the 8-row swizzle atom) address rows s .. s+127 of a TMA-written tile?
// That is what lets one haloed activation tile serve every tap of a dilated conv (tap shift = descriptor row offset).
// Two descriptor variants are tried: base_offset field (bits 49..51) = 0, and = (start_address >> 7) & 7.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I fish_diffusion_b200/csrc tests/native/desc_shift_probe.cu -o build/desc_shift_probe -lcuda
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda.h>
#include <cuda_fp16.h>
#include "fd_tc_ptx.cuh"

constexpr int ROWS = 192;      // rows of the A tile in shared memory
constexpr int NN = 64;         // N of the MMA / rows of W
constexpr int NSHIFT = 40;

template <int BK>
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w, float* out, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int ROWB = BK * 2;
  uint8_t* a_s = smem;                              // ROWS x ROWB (TMA, swizzled)
  uint8_t* w_s = smem + ((ROWS * ROWB + 1023) / 1024) * 1024 + 4096;   // NN x ROWB
  uint64_t* bar = reinterpret_cast<uint64_t*>(w_s + NN * ROWB + 1024);
  uint64_t* mbar = bar + 1;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(mbar, 1); fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, ROWS * ROWB + NN * ROWB);
    // A: box rows limited to 256 -> ROWS=192 fits one box
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(a_s)), "l"(reinterpret_cast<uint64_t>(&tm_a)), "r"(smem_u32(bar)), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(w_s)), "l"(reinterpret_cast<uint64_t>(&tm_w)), "r"(smem_u32(bar)), "r"(0), "r"(0) : "memory");
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  constexpr uint32_t LT = BK == 64 ? 2u : BK == 32 ? 4u : 6u;
  constexpr uint32_t SBO = 8 * ROWB;
  const uint32_t idesc = (1u << 4) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  uint32_t phase = 0;
  for (int s = 0; s < NSHIFT; ++s) {
    if (threadIdx.x == 0) {
      const uint32_t a_addr = smem_u32(a_s) + s * ROWB;
      uint64_t ad = make_kmajor_desc(a_addr, SBO, LT);
      if (mode == 1) ad |= (uint64_t)((a_addr >> 7) & 7) << 49;
      const uint64_t wd = make_kmajor_desc(smem_u32(w_s), SBO, LT);
#pragma unroll
      for (int k = 0; k < BK / 16; ++k) umma_f16(tmem_base, ad + (uint64_t)((k * 32) >> 4), wd + (uint64_t)((k * 32) >> 4), idesc, k != 0);
      umma_commit(mbar);
    }
    mbar_wait(mbar, phase);
    phase ^= 1;
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < NN; c += 16) {
      float v[16];
      tmem_ld16_nowait(taddr + c, v);
      tmem_wait16(v);
      for (int i = 0; i < 16; ++i) out[((size_t)s * 128 + warp * 32 + lane) * NN + c + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
}

template <int BK>
int run() {
  std::vector<__half> a(ROWS * BK), w(NN * BK);
  srand(1 + BK);
  for (auto& x : a) x = __float2half((float)(rand() % 17 - 8) * 0.25f);
  for (auto& x : w) x = __float2half((float)(rand() % 13 - 6) * 0.5f);
  __half *da, *dw; float* dout;
  cudaMalloc(&da, a.size() * 2); cudaMalloc(&dw, w.size() * 2); cudaMalloc(&dout, (size_t)NSHIFT * 128 * NN * 4);
  cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dw, w.data(), w.size() * 2, cudaMemcpyHostToDevice);
  PFN_tmapEncodeTiled enc = get_encode();
  CUtensorMapSwizzle sw = BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : BK == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUtensorMap ta, tw;
  {
    cuuint64_t dims[2] = {(cuuint64_t)BK, (cuuint64_t)ROWS}; cuuint64_t str[1] = {(cuuint64_t)BK * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)ROWS}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&ta, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, da, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode A failed %d\n", (int)r); return 1; }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)BK, (cuuint64_t)NN}; cuuint64_t str[1] = {(cuuint64_t)BK * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)NN}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tw, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, dw, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode W failed %d\n", (int)r); return 1; }
  }
  const int smem = 1024 + ROWS * BK * 2 + 8192 + NN * BK * 2 + 2048;
  cudaFuncSetAttribute(probe_kernel<BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> out((size_t)NSHIFT * 128 * NN);
  for (int mode = 0; mode < 2; ++mode) {
    cudaMemset(dout, 0, out.size() * 4);
    probe_kernel<BK><<<1, 128, smem>>>(ta, tw, dout, mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("BK=%d mode=%d: CUDA error %s\n", BK, mode, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
    printf("BK=%d (swizzle %dB) base_offset mode %d: max |err| per shift:", BK, BK * 2, mode);
    int bad = 0;
    for (int s = 0; s < NSHIFT; ++s) {
      double me = 0;
      for (int r = 0; r < 128; ++r)
        for (int n = 0; n < NN; ++n) {
          double ref = 0;
          for (int k = 0; k < BK; ++k) ref += (double)__half2float(a[(size_t)(r + s) * BK + k]) * (double)__half2float(w[(size_t)n * BK + k]);
          me = fmax(me, fabs(ref - out[((size_t)s * 128 + r) * NN + n]));
        }
      printf(" %d:%.3g", s, me);
      bad += me > 1e-3;
    }
    printf("\n  -> %s (%d of %d shifts wrong)\n", bad ? "MISMATCH" : "ALL SHIFTS EXACT", bad, NSHIFT);
  }
  cudaFree(da); cudaFree(dw); cudaFree(dout);
  return 0;
}

int main() {
  int rc = run<64>();
  rc |= run<32>();
  rc |= run<16>();
  return rc;
}
