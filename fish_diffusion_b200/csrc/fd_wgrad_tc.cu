// Weight-gradient GEMM on tcgen05 straight from channels-last split planes (sm_100a).
//
//   part[s][r][c] = acc_scale * sum_{b in split s} sum_t  ROW[b, t, r] * COL[b, t + shift(c), c]
//
// Time is the contraction axis, and in channels-last storage it is the SLOW axis of both operands.  Instead of
// transposing the activations (the fd_fold_transpose passes of the K-major path), both operands are fed to the tensor
// core as MN-major shared-memory tiles: a TMA box of [64 time steps][64 channels] with the 128-byte swizzle is exactly
// one MN-major SW128 atom column (64 channels contiguous = one 128-byte row per time step, 8 rows per swizzle atom), so
// the instruction descriptor just sets a_major = b_major = MN.  The conv-tap shift of a column segment is a TMA row
// coordinate (zero fill outside [0,T) = the conv zero padding), exactly as in the forward kernel.
//
// Work unit = (split s, 128-row tile, BLOCK_N-column tile); a unit accumulates over its items and all time blocks in
// TMEM and stores one fp32 partial; fd_reduce_batch sums the `splits` partials.  Warp roles and pipelines are those of
// fd_tapgemm_tc.cu (TMA producer / MMA issuer / TMEM allocator / 8 epilogue warps, smem ring + 2 TMEM stages).
#include <cuda.h>
#include <cstring>
#include "fd_common.cuh"
#include "fd_host.h"
#include "fd_tc_ptx.cuh"

namespace {

constexpr int WG_BLOCK_M = 128;
constexpr int WG_BLOCK_K = 64;                 // time steps per pipeline stage
constexpr int WG_EPI_WARP0 = 4;
constexpr int WG_EPI_THREADS = 256;
constexpr int WG_THREADS = WG_EPI_WARP0 * 32 + WG_EPI_THREADS;
constexpr int WG_BOX_BYTES = 64 * 64 * 2;      // one plane of one [64 t][64 ch] box
constexpr int WG_MAX_COL_SEG = 8;

struct FdWgradK {
  int B, T, R, Cc;
  int splits, items_per_split;
  int m_tiles, n_tiles;
  int num_row_seg, num_col_seg;
  int row_src[2], row_coff[2], row_start[2], row_width[2];
  int col_src[WG_MAX_COL_SEG], col_shift[WG_MAX_COL_SEG], col_coff[WG_MAX_COL_SEG], col_start[WG_MAX_COL_SEG],
      col_width[WG_MAX_COL_SEG];
  int row_C[2], col_C[2];
  float* part;
  float acc_scale;
};

// MN-major SW128 operand descriptor (cute::UMMA canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units):
// LBO = byte distance between 64-element atoms along M/N, SBO = byte distance between groups of 8 K rows.
__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}

template <int BLOCK_N, int NPL>
struct WgCfg {
  static constexpr int A_BOXES = WG_BLOCK_M / 64;
  static constexpr int W_BOXES = BLOCK_N / 64;
  static constexpr int A_BYTES = A_BOXES * NPL * WG_BOX_BYTES;
  static constexpr int W_BYTES = W_BOXES * NPL * WG_BOX_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + W_BYTES;
  static constexpr int RAW_STAGES = (192 * 1024) / STAGE_BYTES;
  static constexpr int NUM_STAGES = RAW_STAGES > 8 ? 8 : RAW_STAGES;
  static constexpr int ACC_STAGES = 2;
  static constexpr int TMEM_COLS_RAW = ACC_STAGES * BLOCK_N;
  static constexpr int TMEM_COLS = TMEM_COLS_RAW <= 128 ? 128 : TMEM_COLS_RAW <= 256 ? 256 : 512;
  static constexpr int SCRATCH_BYTES = (WG_EPI_THREADS / 32) * 4096;
  static constexpr int SMEM_BYTES = 1024 + NUM_STAGES * STAGE_BYTES + (2 * NUM_STAGES + 2 * ACC_STAGES) * 8 + 16 +
                                    SCRATCH_BYTES;
};

template <int BLOCK_N, int PREC, int NPL>
__global__ void __launch_bounds__(WG_THREADS, 1)
fd_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_row0, const __grid_constant__ CUtensorMap tm_row1,
                   const __grid_constant__ CUtensorMap tm_col0, const __grid_constant__ CUtensorMap tm_col1,
                   const FdWgradK p) {
  using C = WgCfg<BLOCK_N, NPL>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_base = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::NUM_STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + C::NUM_STAGES;
  uint64_t* tfull_bar = empty_bar + C::NUM_STAGES;
  uint64_t* tempty_bar = tfull_bar + C::ACC_STAGES;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(tempty_bar + C::ACC_STAGES);
  float* scratch_s = reinterpret_cast<float*>(tmem_ptr_s + 4);

  const int warp = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  const int tiles_per_split = p.m_tiles * p.n_tiles;
  const int num_units = p.splits * tiles_per_split;
  const int k_blocks = (p.T + WG_BLOCK_K - 1) / WG_BLOCK_K;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_row0); prefetch_tmap(&tm_row1); prefetch_tmap(&tm_col0); prefetch_tmap(&tm_col1);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::NUM_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < C::ACC_STAGES; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], WG_EPI_THREADS / 32); }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    // =========================================================== TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
        const int s = unit / tiles_per_split;
        const int tile = unit % tiles_per_split;
        const int r0 = (tile / p.n_tiles) * WG_BLOCK_M, c0 = (tile % p.n_tiles) * BLOCK_N;
        // resolve the 64-channel boxes of this tile once: (map, first channel, time shift)
        const CUtensorMap* a_map[C::A_BOXES]; int a_ch[C::A_BOXES];
        const CUtensorMap* w_map[C::W_BOXES]; int w_ch[C::W_BOXES], w_sh[C::W_BOXES];
#pragma unroll
        for (int i = 0; i < C::A_BOXES; ++i) {
          const int r = r0 + 64 * i;
          a_map[i] = &tm_row0; a_ch[i] = p.row_C[0];            // rows past R: a box entirely out of range reads zeros
          for (int g = 0; g < p.num_row_seg; ++g)
            if (r >= p.row_start[g] && r < p.row_start[g] + p.row_width[g]) {
              a_map[i] = p.row_src[g] == 0 ? &tm_row0 : &tm_row1;
              a_ch[i] = p.row_coff[g] + (r - p.row_start[g]);
            }
        }
#pragma unroll
        for (int i = 0; i < C::W_BOXES; ++i) {
          const int c = c0 + 64 * i;
          w_map[i] = &tm_col0; w_ch[i] = p.col_C[0]; w_sh[i] = 0;
          for (int g = 0; g < p.num_col_seg; ++g)
            if (c >= p.col_start[g] && c < p.col_start[g] + p.col_width[g]) {
              w_map[i] = p.col_src[g] == 0 ? &tm_col0 : &tm_col1;
              w_ch[i] = p.col_coff[g] + (c - p.col_start[g]);
              w_sh[i] = p.col_shift[g];
            }
        }
        const int b_end = min(p.B, (s + 1) * p.items_per_split);
        for (int b = s * p.items_per_split; b < b_end; ++b) {
          for (int kb = 0; kb < k_blocks; ++kb) {
            const int t0 = kb * WG_BLOCK_K;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* st = stage_base + stage * C::STAGE_BYTES;
            mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
#pragma unroll
            for (int i = 0; i < C::A_BOXES; ++i)
              tma_load_4d(st + i * NPL * WG_BOX_BYTES, a_map[i], &full_bar[stage], a_ch[i], t0, b, 0);
#pragma unroll
            for (int i = 0; i < C::W_BOXES; ++i)
              tma_load_4d(st + C::A_BYTES + i * NPL * WG_BOX_BYTES, w_map[i], &full_bar[stage], w_ch[i], t0 + w_sh[i],
                          b, 0);
            if (++stage == C::NUM_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================================================== MMA issuer
    if (lane == 0) {
      const uint32_t fmt = PREC == FD_F16 ? 0u : 1u;
      // a_major = b_major = MN (bits 15 / 16): both operands are [time][channel] tiles, channel contiguous
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(WG_BLOCK_M >> 4) << 24);
      constexpr uint32_t LBO = NPL * WG_BOX_BYTES, SBO = 1024;
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
        const int s = unit / tiles_per_split;
        const int n_items = min(p.B, (s + 1) * p.items_per_split) - s * p.items_per_split;
        const int total = n_items * k_blocks;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int it = 0; it < total; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t st = smem_u32(stage_base + stage * C::STAGE_BYTES);
          const uint64_t a_hi = make_mnmajor_desc(st, LBO, SBO);
          const uint64_t a_lo = make_mnmajor_desc(st + WG_BOX_BYTES, LBO, SBO);
          const uint64_t w_hi = make_mnmajor_desc(st + C::A_BYTES, LBO, SBO);
          const uint64_t w_lo = make_mnmajor_desc(st + C::A_BYTES + WG_BOX_BYTES, LBO, SBO);
#pragma unroll
          for (int k = 0; k < WG_BLOCK_K / 16; ++k) {
            const uint64_t adv = (uint64_t)((k * 16 * 128) >> 4);      // 16 time rows of 128 bytes
            if (NPL == 2) {
              umma_f16(d_tmem, a_lo + adv, w_hi + adv, idesc, (it | k) != 0 ? 1u : 0u);
              umma_f16(d_tmem, a_hi + adv, w_lo + adv, idesc, 1u);
              umma_f16(d_tmem, a_hi + adv, w_hi + adv, idesc, 1u);
            } else {
              umma_f16(d_tmem, a_hi + adv, w_hi + adv, idesc, (it | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == C::NUM_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);
        if (++acc == C::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= WG_EPI_WARP0) {
    // =========================================================== epilogue: fp32 partial tile, coalesced rows
    const int q = warp % 4;
    const int half = (warp - WG_EPI_WARP0) / 4;
    constexpr int PER = BLOCK_N / 2;
    const uint32_t my_scratch = smem_u32(scratch_s) + (warp - WG_EPI_WARP0) * 4096;
    const int j4 = (lane & 7) * 4, rsub = lane >> 3;
    int acc = 0; uint32_t acc_phase = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
      const int s = unit / tiles_per_split;
      const int tile = unit % tiles_per_split;
      const int r0 = (tile / p.n_tiles) * WG_BLOCK_M, c0 = (tile % p.n_tiles) * BLOCK_N;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BLOCK_N + ((uint32_t)(q * 32) << 16);
      float* const out = p.part + (size_t)s * p.R * p.Cc;
      for (int c = 0; c < PER; c += 32) {
        float v[32];
        tmem_ld16_nowait(taddr + half * PER + c, *reinterpret_cast<float(*)[16]>(&v[0]));
        tmem_ld16_nowait(taddr + half * PER + c + 16, *reinterpret_cast<float(*)[16]>(&v[16]));
        tmem_wait16(*reinterpret_cast<float(*)[16]>(&v[0]));
        tmem_wait16(*reinterpret_cast<float(*)[16]>(&v[16]));
        float4 a[8];
        warp_transpose_32x32(my_scratch, lane, v, a);
        const int col = c0 + half * PER + c + j4;
#pragma unroll
        for (int pp = 0; pp < 8; ++pp) {
          const int r = r0 + q * 32 + pp * 4 + rsub;
          if (r < p.R)
            *reinterpret_cast<float4*>(out + (size_t)r * p.Cc + col) =
                make_float4(a[pp].x * p.acc_scale, a[pp].y * p.acc_scale, a[pp].z * p.acc_scale, a[pp].w * p.acc_scale);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == C::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
  }
}

// planes [2][B][T][C] as a 4-D tensor (C, T, B, plane); box = [NPL planes][64 t][64 ch], 128-byte swizzle
int make_plane_map(CUtensorMap* m, const uint16_t* ptr, int B, int T, int C, int npl) {
  PFN_tmapEncodeTiled enc = get_encode();
  FD_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)T, (cuuint64_t)B, 2};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)T * C * 2, (cuuint64_t)B * T * C * 2};
  cuuint32_t box[4] = {64, 64, 1, (cuuint32_t)npl};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<uint16_t*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(wgrad) failed: %d (B=%d T=%d C=%d ptr=%p)", (int)r, B, T, C,
             (const void*)ptr);
  return 0;
}


template <int BLOCK_N, int PREC, int NPL>
int launch_wg(const FdWgradK& p, const uint16_t* const* row_ptr, const uint16_t* const* col_ptr, cudaStream_t stream) {
  using C = WgCfg<BLOCK_N, NPL>;
  CUtensorMap tr[2], tc[2];
  for (int i = 0; i < 2; ++i) {
    const int ri = row_ptr[i] != nullptr ? i : 0, ci = col_ptr[i] != nullptr ? i : 0;
    int rc = make_plane_map(&tr[i], row_ptr[ri], p.B, p.T, p.row_C[ri], NPL);
    if (rc) return rc;
    rc = make_plane_map(&tc[i], col_ptr[ci], p.B, p.T, p.col_C[ci], NPL);
    if (rc) return rc;
  }
  auto kern = fd_wgrad_tc_kernel<BLOCK_N, PREC, NPL>;
  static bool attr_set[FD_MAX_DEVICES] = {false};   // the max-dynamic-smem attribute is per device
  const int dev = fd_current_device();
  if (!attr_set[dev]) {
    FD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set[dev] = true;
  }
  const int g_wg_sms = fd_device_sms(dev);
  const int units = p.splits * p.m_tiles * p.n_tiles;
  const int grid = units < g_wg_sms ? units : g_wg_sms;
  kern<<<grid, WG_THREADS, C::SMEM_BYTES, stream>>>(tr[0], tr[1], tc[0], tc[1], p);
  FD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int BLOCK_N>
int launch_wg_prec(const FdWgradK& p, int prec, const uint16_t* const* row_ptr, const uint16_t* const* col_ptr,
                   cudaStream_t stream) {
  const bool single = (prec & FD_SINGLE) != 0;
  if ((prec & 0xF) == FD_F16)
    return single ? launch_wg<BLOCK_N, FD_F16, 1>(p, row_ptr, col_ptr, stream)
                  : launch_wg<BLOCK_N, FD_F16, 2>(p, row_ptr, col_ptr, stream);
  return single ? launch_wg<BLOCK_N, FD_BF16, 1>(p, row_ptr, col_ptr, stream)
                : launch_wg<BLOCK_N, FD_BF16, 2>(p, row_ptr, col_ptr, stream);
}

}  // namespace

extern "C" int fd_wgrad_cl(const fd_wgrad_desc* d, void* stream) {
  FD_DEVICE_GUARD();
  FD_REQUIRE(d != nullptr, "fd_wgrad_cl: null descriptor");
  FD_REQUIRE(d->B > 0 && d->T > 0 && d->splits > 0 && d->splits <= d->B, "fd_wgrad_cl: bad B=%d T=%d splits=%d", d->B,
             d->T, d->splits);
  FD_REQUIRE(d->num_row_seg >= 1 && d->num_row_seg <= 2 && d->num_col_seg >= 1 && d->num_col_seg <= WG_MAX_COL_SEG,
             "fd_wgrad_cl: segment counts out of range (%d rows, %d cols)", d->num_row_seg, d->num_col_seg);
  FD_REQUIRE(d->part != nullptr && d->row_src[0] != nullptr && d->col_src[0] != nullptr, "fd_wgrad_cl: null pointer");
  FdWgradK p;
  memset(&p, 0, sizeof(p));
  p.B = d->B; p.T = d->T; p.splits = d->splits; p.items_per_split = (d->B + d->splits - 1) / d->splits;
  FD_REQUIRE((long long)(p.splits - 1) * p.items_per_split < d->B, "fd_wgrad_cl: splits=%d leaves an empty split", d->splits);
  for (int i = 0; i < 2; ++i) {
    p.row_C[i] = d->row_C[i]; p.col_C[i] = d->col_C[i];
    FD_REQUIRE((d->row_src[i] == nullptr || d->row_C[i] % 8 == 0) && (d->col_src[i] == nullptr || d->col_C[i] % 8 == 0),
               "fd_wgrad_cl: channel counts must be multiples of 8");
  }
  int R = 0, Cc = 0;
  for (int g = 0; g < d->num_row_seg; ++g) {
    const int src = d->row_seg_src[g];
    FD_REQUIRE((src == 0 || src == 1) && d->row_src[src] != nullptr, "fd_wgrad_cl: row segment %d has no source", g);
    FD_REQUIRE(d->row_seg_width[g] > 0 && d->row_seg_width[g] % 64 == 0 && d->row_seg_coff[g] % 8 == 0 &&
                   d->row_seg_coff[g] + d->row_seg_width[g] <= d->row_C[src],
               "fd_wgrad_cl: row segment %d (coff %d width %d) must be a multiple of 64 inside the source", g,
               d->row_seg_coff[g], d->row_seg_width[g]);
    p.row_src[g] = src; p.row_coff[g] = d->row_seg_coff[g]; p.row_start[g] = R; p.row_width[g] = d->row_seg_width[g];
    R += d->row_seg_width[g];
  }
  for (int g = 0; g < d->num_col_seg; ++g) {
    const int src = d->col_seg_src[g];
    FD_REQUIRE((src == 0 || src == 1) && d->col_src[src] != nullptr, "fd_wgrad_cl: column segment %d has no source", g);
    FD_REQUIRE(d->col_seg_width[g] > 0 && d->col_seg_width[g] % 64 == 0 && d->col_seg_coff[g] % 8 == 0 &&
                   d->col_seg_coff[g] + d->col_seg_width[g] <= d->col_C[src],
               "fd_wgrad_cl: column segment %d (coff %d width %d) must be a multiple of 64 inside the source", g,
               d->col_seg_coff[g], d->col_seg_width[g]);
    p.col_src[g] = src; p.col_shift[g] = d->col_seg_shift[g]; p.col_coff[g] = d->col_seg_coff[g];
    p.col_start[g] = Cc; p.col_width[g] = d->col_seg_width[g];
    Cc += d->col_seg_width[g];
  }
  p.R = R; p.Cc = Cc;
  p.num_row_seg = d->num_row_seg; p.num_col_seg = d->num_col_seg;
  p.part = d->part; p.acc_scale = d->acc_scale;
  p.m_tiles = (R + WG_BLOCK_M - 1) / WG_BLOCK_M;
  const int bn = Cc % 256 == 0 ? 256 : Cc % 128 == 0 ? 128 : 64;
  p.n_tiles = Cc / bn;
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (bn == 256) rc = launch_wg_prec<256>(p, d->prec, d->row_src, d->col_src, st);
  else if (bn == 128) rc = launch_wg_prec<128>(p, d->prec, d->row_src, d->col_src, st);
  else rc = launch_wg_prec<64>(p, d->prec, d->row_src, d->col_src, st);
  if (rc == 0) fd_count_launch(1);
  return rc;
}
