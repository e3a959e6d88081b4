// Fused NSF-HiFiGAN ResBlock1 pair on tcgen05 (sm_100a):
//
//     x' = x + c2( lrelu( c1( lrelu(x) ) ) )          (reference models.py:103-110, one iteration of the loop)
//
// in ONE kernel per (c1, c2) pair.  HBM traffic per element: 4 bytes in (split planes of lrelu(x)) + 4 bytes out,
// against 24 for the two separate tap-GEMM launches with an fp32 residual master.
//
//   * the activation tile is loaded ONCE with its halo (128 + (k1-1)*d1 rows, one TMA box per 64-channel block, hi
//     and lo planes in the same box); every conv tap is the SAME shared-memory tile addressed through a tcgen05
//     descriptor whose start address is advanced by whole rows (tests/native/desc_shift_probe.cu: the tensor core
//     swizzles on absolute shared-memory address bits, so any row offset of a TMA-written tile is addressable);
//   * c1's output never leaves the SM: the epilogue warps read the accumulator from TMEM, apply bias + LeakyReLU,
//     split into hi/lo planes and write them into a second swizzled shared-memory tile that is c2's A operand;
//   * the residual x is recovered from the input tile (LeakyReLU is invertible) and pre-loaded, with c2's bias, into
//     c2's TMEM accumulator by tcgen05.st, so GEMM2 accumulates on top of it and the input tile is free for the next
//     tile's TMA load while GEMM2 runs;
//   * three split products per k16 step cost TWO tensor-core instructions: a_hi x [w_hi | w_lo] (the hi and lo weight
//     planes are adjacent rows of one shared-memory tile = one B operand of 2C columns) and a_lo x w_hi; the two
//     accumulator halves are summed in the epilogue.  That cuts the shared-memory operand traffic, which is what
//     bounds narrow-N tcgen05 shapes;
//   * the output goes through a swizzled staging tile and TMA: plane stores, or fp32 store / reduce-add for the
//     multi-receptive-field sum (models.py:426-432).
//
// Warp roles (persistent, one CTA per SM): 0 = weight TMA producer, 1 = MMA issuer, 2 = TMEM allocator,
// 3 = activation-tile TMA producer, 4.. = epilogue (8 warps; 4 at C = 16).
#include <cuda.h>
#include <cstdlib>
#include <cstring>
#include "fd_common.cuh"
#include "fd_host.h"
#include "fd_tc_ptx.cuh"

// epilogue warps: 16 at C = 128 (one block per tile: more warps hide the TMEM / shared-memory latencies of the one
// non-overlapped epilogue; measured 16.1 -> 15.1 ms for k = 11), 8 below (measured slower with 16: 3.13 -> 3.9 ms)
#ifndef FD_RP_EPI_WARPS
#define FD_RP_EPI_WARPS(C) ((C) >= 128 ? 16 : 8)
#endif
// MCAST: the kernel runs as clusters of two CTAs that share ONE weight stream: each CTA loads one plane (hi or lo) of
// every weight unit and multicasts it into both CTAs' rings, and a stage is released when both CTAs' MMA warps have
// committed it.  Halves the L2 -> shared-memory weight traffic at C = 128 (one 128-row block per tile, the pair's
// 1.4 MB of weights per tile at k = 11; ncu: 106.7 GB of TMA loads per launch).  Measured: 1-3 % (15.1 -> 14.9 ms for
// k = 11) -- the weight stream was not the limiter; neither was the depth of the ring (4 / 6 / 8 stages: same time).  What
// was: the per-stage handshake of the single issuing thread (barrier wait, fences, commit) against the 4 MMAs of a 16 KB
// stage -- 32 KB stages (2 x GROUP units, launch_respair) took k = 7 from 11.1 to 9.5 ms on the same box.
#ifndef FD_RP_MCAST
#define FD_RP_MCAST(C) ((C) == 128)
#endif
#ifndef FD_RP_MCAST_DEFAULT
#define FD_RP_MCAST_DEFAULT 1
#endif
#ifndef FD_RP_OCC_DEFAULT
#define FD_RP_OCC_DEFAULT 2
#endif
#ifndef FD_RP_EW_DEFAULT
#define FD_RP_EW_DEFAULT 16
#endif
#ifndef FD_RP_STAGE_MAJOR
#define FD_RP_STAGE_MAJOR(C) ((C) == 64 || (C) == 32)
#endif

namespace {

struct FdResPairK {
  int B, T;
  int k1, d1, k2;
  int h1, h2;        // halos of c1 / c2 in rows
  int r_in, rb, nbox;   // rows of the input tile = nbox TMA boxes of rb rows (rb a multiple of 8)
  int r_out;         // valid output rows per tile = MB*128 - (k2-1)
  int nstages;       // stages of the weight ring (what the input tile of this launch leaves free)
  int group;         // weight units per stage of the ring
  int in_alloc_rows; // rows reserved per plane of the input tile (r_in, or RIN_MAX with the fixed layout)
  int single;        // one product (hi planes only)
  float inv_s1, inv_s2, s2;
  float in_slope_inv, out_slope, planes_scale;
  const float* b1;
  const float* b2;
  // weight units (tap, K slice of BKW channels) that hold non-zero weights, in order, and per listed unit one bit per
  // K16 step (bit i * (BKW/16) + ks): the time-folded C = 16 convs are block-sparse (fd_respair_desc.kmask1/2)
  int n1, n2, masked1, masked2;
  unsigned long long km1, km2;
  unsigned char ul1[64], ul2[64];
};

// A tile is MB blocks of 128 rows with MB * C = 128: every tile holds the same number of elements whatever the
// channel count, so the per-tile latency chain (TMA load -> GEMM1 -> epilogue -> GEMM2 -> epilogue -> TMA store) and
// the (k-1)-row halo are amortised over 1024 rows at C = 16 instead of 128, and the epilogue of block j overlaps the
// MMAs of block j+1.
// OCC = CTAs per SM: 2 halves the tile (MB blocks), the shared memory and the TMEM columns of a CTA, so that the tensor
// phase of one CTA runs under the epilogue phase of the other (the narrow widths spend most of a tile in the epilogues,
// and their accumulators fill all 512 TMEM columns at OCC = 1, which rules out double buffering inside one CTA).
template <int C, int EW = FD_RP_EPI_WARPS(C), int OCC = 1>
struct RpCfg {
  static constexpr int MB = 128 / C / OCC;
  static_assert(MB >= 1, "two CTAs per SM need at least one 128-row block each");
  static constexpr int ROWS = MB * 128;
  static constexpr int BK_A = C >= 64 ? 64 : C;          // channels per shared-memory activation block
  static constexpr int NKB = C / BK_A;
  static constexpr int ROWB = BK_A * 2;                   // bytes per activation row
  static constexpr uint32_t SWZ_A = ROWB == 128 ? 7u : ROWB == 64 ? 3u : 1u;
  static constexpr uint32_t LT_A = ROWB == 128 ? 2u : ROWB == 64 ? 4u : 6u;
  static constexpr int NBOX_MAX = (ROWS + 56 + 255) / 256;
  static constexpr int RIN_MAX = ROWS + 56 + 8 * (NBOX_MAX - 1);
  static constexpr int MID_ROWS = ROWS + 16;              // + zero rows read by the trailing taps of c2
  static constexpr int BKW = C == 128 ? 32 : BK_A;        // K extent of one weight unit
  static constexpr int WROWB = BKW * 2;
  static constexpr uint32_t LT_W = WROWB == 128 ? 2u : WROWB == 64 ? 4u : 6u;
  static constexpr int UNIT_BYTES = 2 * C * WROWB;        // [2 planes][C rows][BKW]
  static constexpr int UNITS_PER_TAP = C / BKW;
  static constexpr int GROUP_RAW = 16384 / UNIT_BYTES;
  static constexpr int GROUP = GROUP_RAW > 8 ? 8 : GROUP_RAW;   // weight units per pipeline stage
  static constexpr int STAGE_BYTES = GROUP * UNIT_BYTES;
  // the input tile is sized per launch (in_plane = r_in * ROWB bytes per plane): what a small halo leaves free becomes
  // extra stages of the weight ring (at C = 128 three stages for k = 11, d = 5 but five for the d = 1 pairs)
  static constexpr int IN_PLANE_BYTES_MAX = RIN_MAX * ROWB;
  static constexpr int MID_PLANE_BYTES = MID_ROWS * ROWB;
  static constexpr int MID_KB_BYTES = 2 * MID_PLANE_BYTES;
  static constexpr int MID_BYTES = NKB * MID_KB_BYTES;     // c1 output; afterwards the staging image of the output
  // STAGE_MAJOR: every weight stage is used by all MB blocks of the tile before it is released, so the pair's weights
  // stream through L2 -> shared memory once per tile instead of once per block (the weight stream is what bounds
  // C = 32 / 64); the price is that the blocks of a tile finish a GEMM together, so their epilogues no longer overlap
  // the MMAs of the following blocks.
  static constexpr bool STAGE_MAJOR = FD_RP_STAGE_MAJOR(C);
  static constexpr bool MCAST = FD_RP_MCAST(C);
  static constexpr int EPI_WARPS = EW;
  static constexpr int HALVES_WANT = C >= 128 ? 4 : C >= 32 ? 2 : 1;
  static constexpr int HALVES = HALVES_WANT > EPI_WARPS / 4 ? EPI_WARPS / 4 : HALVES_WANT;   // column split of a block
  static constexpr int GROUPS = EPI_WARPS / 4 / HALVES;     // warp groups taking alternate blocks
  static constexpr int COLS = C / HALVES;                   // columns per epilogue thread
  static constexpr int EPI_THREADS = EPI_WARPS * 32;
  static constexpr int GTHREADS = EPI_THREADS / GROUPS;
  static constexpr int THREADS = 128 + EPI_THREADS;
  static constexpr int ACC2_OFF = MB * 2 * C;               // acc1[j] at j*2C, acc2[j] at ACC2_OFF + j*2C
  static constexpr int TMEM_COLS = 2 * ACC2_OFF;
  static constexpr int NBAR = 2 + 3 * MB;
  static constexpr int MAX_STAGES = 8;
  static constexpr int HEAD_BYTES = 2048;                   // biases (2C floats <= 1 KB) + barriers, in front of the tiles
  static constexpr int SMEM_BYTES = OCC == 2 ? 113 * 1024 : 227 * 1024;   // always the CTA's whole share: the ring takes what is left
  static constexpr int FIXED_MAX = 1024 + HEAD_BYTES + NKB * 2 * IN_PLANE_BYTES_MAX + MID_BYTES;
  static_assert((SMEM_BYTES - FIXED_MAX) / STAGE_BYTES >= 2, "weight ring needs two stages");
  static_assert(2 * C * 4 + (2 * MAX_STAGES + NBAR + 2) * 8 + 16 <= HEAD_BYTES, "head region too small");
  // stages of the weight ring for an input tile of r_in rows
  static constexpr int stages_for(int r_in, int group) {
    const int n = (SMEM_BYTES - 1024 - HEAD_BYTES - NKB * 2 * r_in * ROWB - MID_BYTES) / (group * UNIT_BYTES);
    return n > MAX_STAGES ? MAX_STAGES : n;
  }
};

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// one plane of a weight unit, written into the same shared-memory offset of BOTH CTAs of the pair; the transaction bytes
// are signalled on the barrier at the same offset in each destination CTA
__device__ __forceinline__ void tma_load_3d_mcast(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                                  uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mcast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

__device__ __forceinline__ uint32_t swz(uint32_t off, uint32_t mask) { return off ^ (((off >> 7) & mask) << 4); }

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(&v[0]);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr) : "memory");
  return r;
}
__device__ __forceinline__ void sts_u4(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <int C, int PREC, bool MC, int EW, int OCC>
__global__ void __launch_bounds__((RpCfg<C, EW, OCC>::THREADS), OCC)
fd_respair_tc_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_w1,
                     const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_out,
                     const __grid_constant__ CUtensorMap tm_out_last, const FdResPairK p) {
  using K = RpCfg<C, EW, OCC>;
  constexpr int MB = K::MB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* bias_s = reinterpret_cast<float*>(smem);                  // b1 [C] | b2 [C]
  uint64_t* w_full = reinterpret_cast<uint64_t*>(bias_s + 2 * C);
  uint64_t* w_empty = w_full + K::MAX_STAGES;
  uint64_t* in_full = w_empty + K::MAX_STAGES;
  const int in_plane = p.in_alloc_rows * K::ROWB;                  // multiple of the swizzle period (rows % 8 == 0)
  const int in_kb = 2 * in_plane;
  const int NUM_STAGES = p.nstages;
  const int GROUP = p.group;                                       // weight units per ring stage
  const int STAGE_BYTES_RT = GROUP * K::UNIT_BYTES;
  uint8_t* in_s = smem + K::HEAD_BYTES;
  uint8_t* mid_s = in_s + K::NKB * in_kb;
  uint8_t* w_s = mid_s + K::MID_BYTES;
  uint64_t* in_empty = in_full + 1;
  uint64_t* acc1_full = in_empty + 1;      // [MB]
  uint64_t* mid_ready = acc1_full + MB;    // [MB]
  uint64_t* acc2_full = mid_ready + MB;    // [MB]
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(acc2_full + MB);

  const int warp = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  const int tiles_t = (p.T + p.r_out - 1) / p.r_out;
  const int num_tiles = p.B * tiles_t;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_in); prefetch_tmap(&tm_w1); prefetch_tmap(&tm_w2); prefetch_tmap(&tm_out); prefetch_tmap(&tm_out_last);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NUM_STAGES; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], MC ? 2 : 1); }
    mbar_init(in_full, 1); mbar_init(in_empty, K::EPI_WARPS);
    for (int j = 0; j < MB; ++j) {
      mbar_init(&acc1_full[j], 1); mbar_init(&mid_ready[j], K::EPI_WARPS / K::GROUPS); mbar_init(&acc2_full[j], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)),
                 "r"((uint32_t)K::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // biases (b2 pre-multiplied by the weight prescale of c2); the 16 trailing rows of the mid tile are zero for the whole kernel
  for (int i = threadIdx.x; i < C; i += blockDim.x) { bias_s[i] = p.b1[i]; bias_s[C + i] = p.b2[i] * p.s2; }
  for (int i = threadIdx.x; i < K::NKB * 2 * K::ROWB; i += blockDim.x)      // 16 rows = ROWB 16-byte chunks per plane
    *reinterpret_cast<uint4*>(mid_s + (i / K::ROWB) * K::MID_PLANE_BYTES + K::ROWS * K::ROWB + (i % K::ROWB) * 16) =
        make_uint4(0, 0, 0, 0);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (MC) cluster_sync_all();      // the peer's barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  const uint32_t acc1 = tmem_base, acc2 = tmem_base + K::ACC2_OFF;
  // every CTA runs the same number of iterations (the pair shares the weight ring); iterations past the last tile only
  // keep the ring turning
  const int iters = (num_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const uint32_t crank = MC ? cluster_ctarank() : 0u;

  const int units1 = p.n1, units2 = p.n2;          // listed (non-zero) weight units of c1 / c2

  if (warp == 0) {
    // =========================================================== weight producer (the pair's weights, once per block)
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        for (int g2 = 0; g2 < 2; ++g2) {
          const CUtensorMap* tm = g2 == 0 ? &tm_w1 : &tm_w2;
          const int units = g2 == 0 ? units1 : units2;
          const unsigned char* ul = g2 == 0 ? p.ul1 : p.ul2;
          const bool pmasked = (g2 == 0 ? p.masked1 : p.masked2) != 0;
          for (int j = 0; j < (K::STAGE_MAJOR ? 1 : MB); ++j) {
            for (int u0 = 0; u0 < units; u0 += GROUP) {
              const int nb = min(GROUP, units - u0);
              mbar_wait(&w_empty[stage], phase ^ 1);
              mbar_expect_tx(&w_full[stage], nb * K::UNIT_BYTES);
              uint8_t* slot = w_s + stage * STAGE_BYTES_RT;
              for (int g = 0; g < nb; ++g) {
                const int u = pmasked ? (int)ul[u0 + g] : u0 + g;
                const int tap = u / K::UNITS_PER_TAP, kw = u % K::UNITS_PER_TAP;
                if (MC)     // this CTA's plane of the unit, into both CTAs
                  tma_load_3d_mcast(slot + g * K::UNIT_BYTES + crank * (C * K::WROWB), tm, &w_full[stage],
                                    tap * C + kw * K::BKW, 0, (int)crank, (uint16_t)3);
                else
                  tma_load_3d(slot + g * K::UNIT_BYTES, tm, &w_full[stage], tap * C + kw * K::BKW, 0, 0);
              }
              if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 3) {
    // =========================================================== activation-tile producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {     // valid tiles only
        const int b = tile / tiles_t, t0 = (tile % tiles_t) * p.r_out;
        mbar_wait(in_empty, (it & 1) ^ 1);
        mbar_expect_tx(in_full, K::NKB * 2 * p.r_in * K::ROWB);
        for (int kb = 0; kb < K::NKB; ++kb)
          for (int pl = 0; pl < 2; ++pl)
            for (int bx = 0; bx < p.nbox; ++bx)
              tma_load_4d(in_s + kb * in_kb + pl * in_plane + bx * p.rb * K::ROWB, &tm_in, in_full,
                          kb * K::BK_A, t0 - p.h2 - p.h1 + bx * p.rb, b, pl);
      }
    }
  } else if (warp == 1) {
    // =========================================================== MMA issuer
    // The WHOLE warp runs the loops (every value is warp-uniform, so descriptors and addresses stay in uniform
    // registers); only the tcgen05 instructions are issued by one elected lane.  With the loops inside `if (lane == 0)`
    // the compiler wraps every UTCHMMA in an ELECT / R2UR.BROADCAST / branch sequence (~100 cycles per instruction),
    // which is what bounds the narrow-N shapes (an N = 16..64 instruction occupies the tensor core for 8..32 cycles).
    const uint32_t fmt = PREC == FD_F16 ? 0u : 1u;
    const uint32_t idesc_base = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc_2c = idesc_base | ((uint32_t)((2 * C) >> 3) << 17);
    const uint32_t idesc_c = idesc_base | ((uint32_t)(C >> 3) << 17);
    constexpr uint32_t SBO_A = 8 * K::ROWB, SBO_W = 8 * K::WROWB;
    constexpr uint64_t DESC_HI_A = ((uint64_t)((SBO_A >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)K::LT_A << 61) | ((uint64_t)1 << 16);
    constexpr uint64_t DESC_HI_W = ((uint64_t)((SBO_W >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)K::LT_W << 61) | ((uint64_t)1 << 16);
    const bool single = p.single != 0;
    int stage = 0; uint32_t phase = 0;
    for (uint32_t it = 0; it < (uint32_t)iters; ++it) {
      const bool valid = (int)blockIdx.x + (int)it * (int)gridDim.x < num_tiles;
      if (!valid) {     // no tile left for this CTA: consume and release the shared weight stages only
        const int stages_per_tile = ((units1 + GROUP - 1) / GROUP + (units2 + GROUP - 1) / GROUP) *
                                    (K::STAGE_MAJOR ? 1 : MB);
        for (int sidx = 0; sidx < stages_per_tile; ++sidx) {
          mbar_wait(&w_full[stage], phase);
          if (elect_one()) { if (MC) umma_commit_mcast(&w_empty[stage], 3); else umma_commit(&w_empty[stage]); }
          __syncwarp();
          if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
        }
        continue;
      }
#pragma unroll 1
      for (int g2 = 0; g2 < 2; ++g2) {
        const int units = g2 == 0 ? units1 : units2;
        const unsigned char* ul = g2 == 0 ? p.ul1 : p.ul2;
        const unsigned long long km = g2 == 0 ? p.km1 : p.km2;
        const bool masked = (g2 == 0 ? p.masked1 : p.masked2) != 0;
        constexpr int KS = K::BKW / 16;
        const uint32_t a_base = g2 == 0 ? smem_u32(in_s) : smem_u32(mid_s);
        const uint32_t a_kb = g2 == 0 ? in_kb : K::MID_KB_BYTES;
        const uint32_t a_plane = g2 == 0 ? (uint32_t)in_plane : (uint32_t)K::MID_PLANE_BYTES;
        const uint32_t tap_bytes = (uint32_t)(g2 == 0 ? p.d1 : 1) * K::ROWB;
        if (g2 == 0) { mbar_wait(in_full, it & 1); tc_fence_after(); }
        if (K::STAGE_MAJOR) {
          if (g2 == 1) {   // every block's mid rows and residual-initialised accumulator
#pragma unroll 1
            for (int j = 0; j < MB; ++j) mbar_wait(&mid_ready[j], it & 1);
            tc_fence_after();
          }
          int ui = 0;
          uint32_t started = g2;                   // GEMM2 accumulates on the residual-initialised accumulator
#pragma unroll 1
          for (int u0 = 0; u0 < units; u0 += GROUP) {
            const int nb = min(GROUP, units - u0);
            mbar_wait(&w_full[stage], phase);
            tc_fence_after();
            const uint32_t w_stage = smem_u32(w_s + stage * STAGE_BYTES_RT);
#pragma unroll 1
            for (int g = 0; g < nb; ++g, ++ui) {
              const int u = masked ? (int)ul[ui] : ui;      // dense convs: no table look-up on the issue path
              const int tap = u / K::UNITS_PER_TAP, kw = u % K::UNITS_PER_TAP;
              const int ch = kw * K::BKW;
              const uint32_t a0 = a_base + (ch / K::BK_A) * a_kb + (uint32_t)tap * tap_bytes + (ch % K::BK_A) * 2;
              const uint64_t dw = DESC_HI_W | (uint64_t)(((w_stage + g * K::UNIT_BYTES) & 0x3FFFF) >> 4);
              const uint32_t kbits = masked ? (uint32_t)(km >> (ui * KS)) & ((1u << KS) - 1u) : ((1u << KS) - 1u);
              const uint32_t first = started;
              if (elect_one()) {
#pragma unroll
                for (int j = 0; j < MB; ++j) {
                  const uint32_t d_tmem = (g2 == 0 ? acc1 : acc2) + j * 2 * C;
                  const uint32_t a_hi = a0 + (uint32_t)(j * 128) * K::ROWB;
                  const uint64_t da_hi = DESC_HI_A | (uint64_t)((a_hi & 0x3FFFF) >> 4);
                  const uint64_t da_lo = DESC_HI_A | (uint64_t)(((a_hi + a_plane) & 0x3FFFF) >> 4);
                  uint32_t accum = first;
#pragma unroll
                  for (int k = 0; k < K::BKW / 16; ++k) {
                    if (!((kbits >> k) & 1u)) continue;             // an all-zero K16 slice of this tap
                    if (!single) {
                      umma_f16(d_tmem, da_hi + 2 * k, dw + 2 * k, idesc_2c, accum);
                      umma_f16(d_tmem, da_lo + 2 * k, dw + 2 * k, idesc_c, 1u);
                    } else {
                      umma_f16(d_tmem, da_hi + 2 * k, dw + 2 * k, idesc_c, accum);
                    }
                    accum = 1u;
                  }
                }
              }
              if (kbits) started = 1u;
              __syncwarp();
            }
            if (elect_one()) { if (MC) umma_commit_mcast(&w_empty[stage], 3); else umma_commit(&w_empty[stage]); }
            __syncwarp();
            if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
          }
          if (elect_one()) {
            for (int j = 0; j < MB; ++j) umma_commit(g2 == 0 ? &acc1_full[j] : &acc2_full[j]);
          }
          __syncwarp();
        } else {
#pragma unroll 1
        for (int j = 0; j < MB; ++j) {
          const uint32_t d_tmem = (g2 == 0 ? acc1 : acc2) + j * 2 * C;
          if (g2 == 1) {   // GEMM2 of block j reads mid rows of blocks j and j+1 (and the residual-initialised accumulator)
            mbar_wait(&mid_ready[j + 1 < MB ? j + 1 : j], it & 1);
            tc_fence_after();
          }
          int ui = 0;
          uint32_t started = g2;
#pragma unroll 1
          for (int u0 = 0; u0 < units; u0 += GROUP) {
            const int nb = min(GROUP, units - u0);
            mbar_wait(&w_full[stage], phase);
            tc_fence_after();
            const uint32_t w_stage = smem_u32(w_s + stage * STAGE_BYTES_RT);
#pragma unroll 1
            for (int g = 0; g < nb; ++g, ++ui) {
              const int u = masked ? (int)ul[ui] : ui;      // dense convs: no table look-up on the issue path
              const int tap = u / K::UNITS_PER_TAP, kw = u % K::UNITS_PER_TAP;
              const int ch = kw * K::BKW;
              const uint32_t kbits = masked ? (uint32_t)(km >> (ui * KS)) & ((1u << KS) - 1u) : ((1u << KS) - 1u);
              const uint32_t a_hi = a_base + (ch / K::BK_A) * a_kb + (uint32_t)(j * 128) * K::ROWB + (uint32_t)tap * tap_bytes +
                                    (ch % K::BK_A) * 2;
              const uint64_t da_hi = DESC_HI_A | (uint64_t)((a_hi & 0x3FFFF) >> 4);
              const uint64_t da_lo = DESC_HI_A | (uint64_t)(((a_hi + a_plane) & 0x3FFFF) >> 4);
              const uint64_t dw = DESC_HI_W | (uint64_t)(((w_stage + g * K::UNIT_BYTES) & 0x3FFFF) >> 4);
              if (elect_one()) {
                uint32_t accum = started;
#pragma unroll
                for (int k = 0; k < K::BKW / 16; ++k) {
                  if (!((kbits >> k) & 1u)) continue;
                  if (!single) {
                    umma_f16(d_tmem, da_hi + 2 * k, dw + 2 * k, idesc_2c, accum);     // a_hi x [w_hi | w_lo] -> [0,2C)
                    umma_f16(d_tmem, da_lo + 2 * k, dw + 2 * k, idesc_c, 1u);         // a_lo x w_hi         -> [0,C)
                  } else {
                    umma_f16(d_tmem, da_hi + 2 * k, dw + 2 * k, idesc_c, accum);
                  }
                  accum = 1u;
                }
              }
              if (kbits) started = 1u;
              __syncwarp();
            }
            if (elect_one()) { if (MC) umma_commit_mcast(&w_empty[stage], 3); else umma_commit(&w_empty[stage]); }
            __syncwarp();
            if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
          }
          if (elect_one()) umma_commit(g2 == 0 ? &acc1_full[j] : &acc2_full[j]);
          __syncwarp();
        }
        }
      }
    }
  } else if (warp >= 4) {
    // =========================================================== epilogue
    const int q = warp % 4;
    const int wg = (warp - 4) / 4;
    const int grp = wg / K::HALVES;                   // which blocks this warp takes (j % GROUPS == grp)
    const int half = wg % K::HALVES;                  // which columns of a block
    const int row = q * 32 + lane;
    const int col_base = half * K::COLS;
    const int etid = threadIdx.x - 128;
    const bool issuer = (etid % K::GTHREADS) == 0;    // first thread of a group issues that group's TMA stores
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t in_u = smem_u32(in_s), mid_u = smem_u32(mid_s);
    const float slope_mid = 0.1f;
    const float s2_neg = p.s2 * p.in_slope_inv;
    const float out_c = p.inv_s2 * p.planes_scale, out_cs = out_c * p.out_slope;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int b = tile / tiles_t, t0 = (tile % tiles_t) * p.r_out;
      // ---- phase 1a: residual + c2 bias -> accumulator 2 (pre-scaled by the weight prescale of c2), zeros in the [w_lo]
      // half.  Needs only the input tile (and this thread's own reads of accumulator 2 for the previous tile, which are
      // complete), so it runs while GEMM1 of this tile is still being issued.
      mbar_wait(in_full, it & 1);                    // visibility of the TMA-written input tile to these threads
#pragma unroll 1
      for (int j = grp; j < MB; j += K::GROUPS) {
        const int n = j * 128 + row + p.h1 + p.h2;
#pragma unroll
        for (int c16 = 0; c16 < K::COLS / 16; ++c16) {
          const int col = col_base + c16 * 16;
          const int kb = col / K::BK_A, cc = col % K::BK_A;
          const uint32_t base = in_u + kb * in_kb;
          float v[16];
#pragma unroll
          for (int hq = 0; hq < 2; ++hq) {
            const uint32_t off = swz((uint32_t)n * K::ROWB + (uint32_t)(cc / 8 + hq) * 16, K::SWZ_A);
            const uint4 h4 = lds_u4(base + off), l4 = lds_u4(base + in_plane + off);
            const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              float a0, a1;
              fd_combine2(hw[jj], lw[jj], PREC, a0, a1);
              // x = p >= 0 ? p : p / slope (the planes hold lrelu(x)); (x + b2) * s2 as one fma
              v[hq * 8 + 2 * jj] = fmaf(a0, a0 >= 0.f ? p.s2 : s2_neg, bias_s[C + col + hq * 8 + 2 * jj]);
              v[hq * 8 + 2 * jj + 1] = fmaf(a1, a1 >= 0.f ? p.s2 : s2_neg, bias_s[C + col + hq * 8 + 2 * jj + 1]);
            }
          }
          tmem_st16(acc2 + j * 2 * C + lane_addr + col, v);
          if (!p.single) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = 0.f;
            tmem_st16(acc2 + j * 2 * C + lane_addr + C + col, v);
          }
        }
      }
      tmem_st_wait();
      // ---- phase 1b per block: GEMM1 done -> bias, LeakyReLU, zero outside [0,T), split planes -> mid tile (c2's A operand)
#pragma unroll 1
      for (int j = grp; j < MB; j += K::GROUPS) {
        mbar_wait(&acc1_full[j], it & 1);
        tc_fence_after();
        if (j == grp) {
          // the staging image (= mid tile) of the previous tile must have been read by its TMA stores
          if (issuer) bulk_wait_read0();
          asm volatile("bar.sync 8, %0;" ::"r"(K::EPI_THREADS) : "memory");
        }
        const int m = j * 128 + row;                   // row of the mid tile
        const int t = t0 - p.h2 + m;
        const bool ok = t >= 0 && t < p.T;
        const bool all_ok = __all_sync(0xffffffffu, ok);
#pragma unroll
        for (int c16 = 0; c16 < K::COLS / 16; ++c16) {
          const int col = col_base + c16 * 16;
          float a[16], a2[16];
          tmem_ld16_nowait(acc1 + j * 2 * C + lane_addr + col, a);
          if (!p.single) tmem_ld16_nowait(acc1 + j * 2 * C + lane_addr + C + col, a2);
          tmem_wait16(a);
          if (!p.single) {
            tmem_wait16(a2);
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] += a2[i];
          }
          uint32_t hi[8], lo[8];
          if (all_ok) {
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              const float y0 = fmaf(a[2 * jj], p.inv_s1, bias_s[col + 2 * jj]);
              const float y1 = fmaf(a[2 * jj + 1], p.inv_s1, bias_s[col + 2 * jj + 1]);
              fd_split2(fmaxf(y0, slope_mid * y0), fmaxf(y1, slope_mid * y1), PREC, hi[jj], lo[jj]);
            }
          } else {
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              const float y0 = fmaf(a[2 * jj], p.inv_s1, bias_s[col + 2 * jj]);
              const float y1 = fmaf(a[2 * jj + 1], p.inv_s1, bias_s[col + 2 * jj + 1]);
              fd_split2(ok ? fmaxf(y0, slope_mid * y0) : 0.f, ok ? fmaxf(y1, slope_mid * y1) : 0.f, PREC, hi[jj], lo[jj]);
            }
          }
          const int kb = col / K::BK_A, cc = col % K::BK_A;
          const uint32_t base = mid_u + kb * K::MID_KB_BYTES;
#pragma unroll
          for (int hq = 0; hq < 2; ++hq) {
            const uint32_t off = swz((uint32_t)m * K::ROWB + (uint32_t)(cc / 8 + hq) * 16, K::SWZ_A);
            sts_u4(base + off, hi[hq * 4], hi[hq * 4 + 1], hi[hq * 4 + 2], hi[hq * 4 + 3]);
            sts_u4(base + K::MID_PLANE_BYTES + off, lo[hq * 4], lo[hq * 4 + 1], lo[hq * 4 + 2], lo[hq * 4 + 3]);
          }
        }
        fence_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&mid_ready[j]);
          if (j + K::GROUPS >= MB) mbar_arrive(in_empty);      // GEMM1 of every block has read the input tile, and so has this warp
        }
      }

      // ---- phase 2 per block: GEMM2 done (accumulator 2 = (x + b2) * s2 + conv products) -> output planes via TMA
#pragma unroll 1
      for (int j = grp; j < MB; j += K::GROUPS) {
        mbar_wait(&acc2_full[j], it & 1);
        tc_fence_after();
        const int m = j * 128 + row;
#pragma unroll
        for (int c16 = 0; c16 < K::COLS / 16; ++c16) {
          const int col = col_base + c16 * 16;
          float a[16], a2[16];
          tmem_ld16_nowait(acc2 + j * 2 * C + lane_addr + col, a);
          if (!p.single) tmem_ld16_nowait(acc2 + j * 2 * C + lane_addr + C + col, a2);
          tmem_wait16(a);
          if (!p.single) {
            tmem_wait16(a2);
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] += a2[i];
          }
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj)      // lrelu(a * inv_s2) * planes_scale = max(a * c, a * c * slope), 0 <= slope <= 1
            fd_split2(fmaxf(a[2 * jj] * out_c, a[2 * jj] * out_cs), fmaxf(a[2 * jj + 1] * out_c, a[2 * jj + 1] * out_cs), PREC,
                      hi[jj], lo[jj]);
          const int kb = col / K::BK_A, cc = col % K::BK_A;
          const uint32_t base = mid_u + kb * K::MID_KB_BYTES;
#pragma unroll
          for (int hq = 0; hq < 2; ++hq) {
            const uint32_t off = swz((uint32_t)m * K::ROWB + (uint32_t)(cc / 8 + hq) * 16, K::SWZ_A);
            sts_u4(base + off, hi[hq * 4], hi[hq * 4 + 1], hi[hq * 4 + 2], hi[hq * 4 + 3]);
            sts_u4(base + K::MID_PLANE_BYTES + off, lo[hq * 4], lo[hq * 4 + 1], lo[hq * 4 + 2], lo[hq * 4 + 3]);
          }
        }
        fence_async_smem();
        tc_fence_before();
        asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(K::GTHREADS) : "memory");
        if (issuer) {
          const CUtensorMap* tm = j == MB - 1 ? &tm_out_last : &tm_out;
          for (int kb = 0; kb < K::NKB; ++kb)
            for (int pl = 0; pl < 2; ++pl)
              tma_store_4d(tm, mid_u + kb * K::MID_KB_BYTES + pl * K::MID_PLANE_BYTES + (uint32_t)(j * 128) * K::ROWB,
                           kb * K::BK_A, t0 + j * 128, b, pl);
          bulk_commit();
        }
      }
    }
    if (issuer) bulk_wait0();
  }

  // ---------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (MC) cluster_sync_all();      // no multicast write / remote arrive may target a CTA that has exited
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)K::TMEM_COLS)
                 : "memory");
  }
}

// ------------------------------------------------------------------ host side
CUtensorMapSwizzle swizzle_for_bytes(int row_bytes) {
  return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                                         : CU_TENSOR_MAP_SWIZZLE_32B;
}

// planes [2][B][T][C] (uint16): box {bk, rows, 1, nplanes}
int make_planes_map(CUtensorMap* m, const uint16_t* ptr, int B, int T, int C, int bk, int rows, int nplanes) {
  PFN_tmapEncodeTiled enc = get_encode();
  FD_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)T, (cuuint64_t)B, 2};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)T * C * 2, (cuuint64_t)B * T * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)bk, (cuuint32_t)rows, 1, (cuuint32_t)nplanes};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<uint16_t*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(bk * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(planes) failed: %d (B=%d T=%d C=%d bk=%d rows=%d)", (int)r, B, T,
             C, bk, rows);
  return 0;
}

// packed weights [2][C][K] (uint16): box {bkw, C, planes} (both planes of a unit, or one per CTA of a multicast pair)
int make_wpair_map(CUtensorMap* m, const uint16_t* ptr, int C, int Ktot, int bkw, int planes) {
  PFN_tmapEncodeTiled enc = get_encode();
  FD_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {(cuuint64_t)Ktot, (cuuint64_t)C, 2};
  cuuint64_t strides[2] = {(cuuint64_t)Ktot * 2, (cuuint64_t)C * Ktot * 2};
  cuuint32_t box[3] = {(cuuint32_t)bkw, (cuuint32_t)C, (cuuint32_t)planes};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<uint16_t*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(bkw * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(w pair) failed: %d (C=%d K=%d)", (int)r, C, Ktot);
  return 0;
}

template <int C, int PREC, bool MC, int EW, int OCC>
int launch_respair(const fd_respair_desc& d, FdResPairK p, cudaStream_t stream) {
  using K = RpCfg<C, EW, OCC>;
  // tile geometry: MB blocks of 128 rows; the input tile is nbox TMA boxes of rb rows
  const int need = K::ROWS + (p.k1 - 1) * p.d1;
  p.nbox = (need + 255) / 256;
  p.rb = ((need + p.nbox - 1) / p.nbox + 7) / 8 * 8;
  p.r_in = p.nbox * p.rb;
  p.r_out = K::ROWS - (p.k2 - 1);
  FD_REQUIRE(p.r_in <= K::RIN_MAX && p.rb <= 256, "fd_respair_fwd: input tile of %d rows exceeds the shared-memory tile", p.r_in);
  {
    const char* e = getenv("FD_RP_FIXED_TILE");   // experiment knob: 1 = reserve RIN_MAX rows whatever the halo
    p.in_alloc_rows = (e && e[0] == '1') ? K::RIN_MAX : p.r_in;
    // a stage of the ring holds 2 x GROUP weight units (32 KB) when at least two such stages fit: the per-stage
    // handshake (barrier wait, fences, commit) of the single issuing thread is what the MMAs of a 16 KB stage cannot hide
    const char* gm = getenv("FD_RP_GROUP_MULT");
    const int mult = gm ? atoi(gm) : 2;
    p.group = K::GROUP;
    if (mult >= 2 && K::stages_for(p.in_alloc_rows, mult * K::GROUP) >= 2) p.group = mult * K::GROUP;
    p.nstages = K::stages_for(p.in_alloc_rows, p.group);
    const char* m = getenv("FD_RP_MAX_STAGES");
    if (m && atoi(m) >= 2 && atoi(m) < p.nstages) p.nstages = atoi(m);
  }
  CUtensorMap tin, tw1, tw2, tout, tout_last;
  int rc = make_planes_map(&tin, d.in_planes, p.B, p.T, C, K::BK_A, p.rb, 1);
  if (rc) return rc;
  rc = make_wpair_map(&tw1, d.w1, C, p.k1 * C, K::BKW, MC ? 1 : 2);
  if (rc) return rc;
  rc = make_wpair_map(&tw2, d.w2, C, p.k2 * C, K::BKW, MC ? 1 : 2);
  if (rc) return rc;
  rc = make_planes_map(&tout, d.out_planes, p.B, p.T, C, K::BK_A, 128, 1);
  if (rc) return rc;
  rc = make_planes_map(&tout_last, d.out_planes, p.B, p.T, C, K::BK_A, 128 - (p.k2 - 1), 1);
  if (rc) return rc;
  auto kern = fd_respair_tc_kernel<C, PREC, MC, EW, OCC>;
  static bool attr_set[FD_MAX_DEVICES] = {false};
  const int dev = fd_current_device();
  if (!attr_set[dev]) {
    FD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM_BYTES));
    FD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr_set[dev] = true;
  }
  const int tiles = p.B * ((p.T + p.r_out - 1) / p.r_out);
  const int sms = fd_device_sms(dev) * OCC;
  int grid = tiles < sms ? tiles : sms;
  fd_prof_begin(C == 128 ? 12 : C == 64 ? 13 : C == 32 ? 14 : 15, stream);
  if (MC) {
    grid = (grid + 1) / 2 * 2;           // whole pairs; a CTA without tiles only keeps the shared weight ring turning
    if (grid > sms) grid -= 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(K::THREADS); cfg.dynamicSmemBytes = K::SMEM_BYTES; cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    FD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tin, tw1, tw2, tout, tout_last, p));
  } else {
    kern<<<grid, K::THREADS, K::SMEM_BYTES, stream>>>(tin, tw1, tw2, tout, tout_last, p);
  }
  fd_prof_end(stream);
  FD_CHECK_CUDA(cudaGetLastError());
  fd_count_launch(1);
  return 0;
}

template <int C, bool MC, int EW, int OCC>
int launch_respair_mc(const fd_respair_desc& d, const FdResPairK& p, cudaStream_t stream) {
  return (d.prec & 0xF) == FD_F16 ? launch_respair<C, FD_F16, MC, EW, OCC>(d, p, stream)
                                  : launch_respair<C, FD_BF16, MC, EW, OCC>(d, p, stream);
}

template <int C>
int launch_respair_prec(const fd_respair_desc& d, const FdResPairK& p, cudaStream_t stream) {
  // FD_RP_MCAST(C) names the widths that have the 2-CTA weight-multicast variant; FD_RP_MCAST_ON=0/1 picks at run time
  static const int mc_env = [] { const char* e = getenv("FD_RP_MCAST_ON"); return e ? atoi(e) : -1; }();
  // epilogue warps of the stage-major widths (C = 32 / 64): FD_RP_EW=8/16; CTAs per SM at C = 32: FD_RP_OCC=1/2
  static const int ew_env = [] { const char* e = getenv("FD_RP_EW"); return e ? atoi(e) : FD_RP_EW_DEFAULT; }();
  static const int occ_env = [] { const char* e = getenv("FD_RP_OCC"); return e ? atoi(e) : FD_RP_OCC_DEFAULT; }();
  if constexpr (RpCfg<C>::MCAST) {
    if (mc_env != 0 && (mc_env == 1 || FD_RP_MCAST_DEFAULT)) return launch_respair_mc<C, true, FD_RP_EPI_WARPS(C), 1>(d, p, stream);
  }
  if constexpr (C == 32) {
    if (occ_env == 2) return launch_respair_mc<C, false, 8, 2>(d, p, stream);
  }
  if constexpr (C == 32 || C == 64) {
    if (ew_env == 16) return launch_respair_mc<C, false, 16, 1>(d, p, stream);
    return launch_respair_mc<C, false, 8, 1>(d, p, stream);
  }
  return launch_respair_mc<C, false, FD_RP_EPI_WARPS(C), 1>(d, p, stream);
}

}  // namespace

extern "C" int fd_respair_supported(int C, int k1, int d1, int k2) {
  if (C != 16 && C != 32 && C != 64 && C != 128) return 0;
  if (k1 < 1 || k2 < 1 || (k1 & 1) == 0 || (k2 & 1) == 0 || d1 < 1) return 0;
  if (k2 - 1 > 16) return 0;                             // zero rows behind the mid tile
  if ((k1 - 1) * d1 > 56) return 0;                      // halo rows of the input tile
  if ((k2 - 1) / 2 > (k1 - 1) / 2 * d1) return 0;       // the residual rows must lie inside the input tile
  return 1;
}

extern "C" int fd_respair_fwd(const fd_respair_desc* d, void* stream) {
  FD_DEVICE_GUARD();
  FD_REQUIRE(d != nullptr, "fd_respair_fwd: null descriptor");
  FD_REQUIRE(fd_respair_supported(d->C, d->k1, d->d1, d->k2), "fd_respair_fwd: unsupported shape C=%d k1=%d d1=%d k2=%d",
             d->C, d->k1, d->d1, d->k2);
  FD_REQUIRE(d->B > 0 && d->T > 0, "fd_respair_fwd: bad shape B=%d T=%d", d->B, d->T);
  FD_REQUIRE(d->in_planes && d->w1 && d->w2 && d->b1 && d->b2 && d->out_planes, "fd_respair_fwd: null pointer");
  FD_REQUIRE(d->in_slope > 0.f && d->in_slope <= 1.f, "fd_respair_fwd: the input LeakyReLU slope must be in (0, 1] (it is inverted)");
  FD_REQUIRE(d->out_slope >= 0.f && d->out_slope <= 1.f, "fd_respair_fwd: the output LeakyReLU slope must be in [0, 1]");
  FD_REQUIRE((const void*)d->out_planes != (const void*)d->in_planes, "fd_respair_fwd: in-place is not supported (halo reads)");
  FdResPairK p;
  memset(&p, 0, sizeof(p));
  p.B = d->B; p.T = d->T; p.k1 = d->k1; p.d1 = d->d1; p.k2 = d->k2;
  p.h1 = (d->k1 - 1) / 2 * d->d1; p.h2 = (d->k2 - 1) / 2;
  p.single = (d->prec & FD_SINGLE) ? 1 : 0;
  p.inv_s1 = d->w1_inv_scale; p.inv_s2 = d->w2_inv_scale; p.s2 = 1.f / d->w2_inv_scale;
  p.in_slope_inv = 1.f / d->in_slope; p.out_slope = d->out_slope; p.planes_scale = d->planes_scale;
  p.b1 = d->b1; p.b2 = d->b2;
  {
    // weight units of BKW channels per tap (BKW as in RpCfg: 32 at C = 128, else min(C, 64)) and their K16 steps
    const int bkw = d->C == 128 ? 32 : (d->C >= 64 ? 64 : d->C), upt = d->C / bkw, ks = bkw / 16, kst = d->C / 16;
    FD_REQUIRE(d->k1 * upt <= 64 && d->k2 * upt <= 64, "fd_respair_fwd: more than 64 weight units (k1=%d k2=%d C=%d)", d->k1,
               d->k2, d->C);
    auto fill = [&](int k, unsigned long long kmask, unsigned char* ul, int& n, unsigned long long& km) {
      n = 0; km = 0;
      const bool masked = kmask != 0ull && k * kst <= 64;
      if (masked) {
        for (int u = 0; u < k * upt; ++u) {
          const unsigned bits = (unsigned)(kmask >> ((u / upt) * kst + (u % upt) * ks)) & ((1u << ks) - 1u);
          if (bits) { km |= (unsigned long long)bits << (n * ks); ul[n++] = (unsigned char)u; }
        }
      }
      if (n == 0) {                    // dense (or a mask without any set bit inside the conv)
        for (int u = 0; u < k * upt; ++u) ul[u] = (unsigned char)u;
        n = k * upt; km = 0;
        return false;
      }
      return true;
    };
    const bool m1 = fill(d->k1, d->kmask1, p.ul1, p.n1, p.km1), m2 = fill(d->k2, d->kmask2, p.ul2, p.n2, p.km2);
    p.masked1 = m1 ? 1 : 0; p.masked2 = m2 ? 1 : 0;
  }
  cudaStream_t st = (cudaStream_t)stream;
  switch (d->C) {
    case 128: return launch_respair_prec<128>(*d, p, st);
    case 64: return launch_respair_prec<64>(*d, p, st);
    case 32: return launch_respair_prec<32>(*d, p, st);
    default: return launch_respair_prec<16>(*d, p, st);
  }
}
