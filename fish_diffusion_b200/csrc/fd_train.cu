// Memory-bound kernels of the WaveNet training step (backward pass).  The six gradient GEMMs per residual block
// are tap-GEMM launches (fd_tapgemm_*.cu): data gradients use transposed packed weights with mirrored tap shifts,
// weight gradients use the same kernel with "rows" = output channels and K = time, fed by the folded transposes
// produced here (wavenet.py:106-120 differentiated by hand; checked against the reference's autograd in the tests).
#include <cstring>
#include "fd_common.cuh"
#include "fd_host.h"

namespace {

// planes [2][B][T][C] (+ optional per-(item, channel) addend d) -> planes [2][C][B][Tp], item b's T samples start at
// column PAD of its Tp-wide span, everything else is zero.  `mode`: 0 plain, 1 z = sigmoid(g)*tanh(f) computed from a
// packed pre-activation tensor (C = residual channels, source has 2C packed columns), 2 relu-mask (src is a fp32
// gradient [B][T][C], `aux` planes give the forward activation; value = grad * (act > 0)).
template <int MODE>
__global__ void k_fold_transpose(const uint16_t* __restrict__ src, const float* __restrict__ src_f32,
                                 const uint16_t* __restrict__ aux, const float* __restrict__ addvec, int add_bstride,
                                 uint16_t* __restrict__ dst, int B, int T, int C, int Tp, int pad, float scale,
                                 int gate_tile, int prec, int dst_rows, int dst_row0) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;   // t0 indexes the padded axis
  const size_t splane = MODE == 1 ? (size_t)B * T * 2 * C : (size_t)B * T * C;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int tp = t0 + i, c = c0 + threadIdx.x;
    const int t = tp - pad;
    float v = 0.f;
    if (t >= 0 && t < T && c < C) {
      const size_t row = (size_t)b * T + t;
      if (MODE == 0) {
        const size_t off = row * C + c;
        v = fd_combine(src[off], src[splane + off], prec);
        if (addvec != nullptr) v += addvec[(size_t)b * add_bstride + c];
      } else if (MODE == 1) {
        const int half = gate_tile / 2;
        const int ng = (c / half) * gate_tile + (c % half);
        const size_t off = row * 2 * C + ng;
        const float g = fd_combine(src[off], src[splane + off], prec);
        const float f = fd_combine(src[off + half], src[splane + off + half], prec);
        v = fd_sigmoid(g) * fd_tanh(f);
      } else {
        const size_t off = row * C + c;
        const float a = fd_combine(aux[off], aux[splane + off], prec);
        v = a > 0.f ? src_f32[off] : 0.f;
      }
      v *= scale;
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  const size_t dplane = (size_t)dst_rows * B * Tp;      // the destination may be a taller stacked operand
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, tp = t0 + threadIdx.x;
    if (c < C && tp < Tp) {
      uint16_t hi, lo;
      fd_split(tile[threadIdx.x][i], prec, hi, lo);
      const size_t off = ((size_t)(dst_row0 + c) * B + b) * Tp + tp;
      dst[off] = hi;
      dst[dplane + off] = lo;
    }
  }
}

// dz (fp32 [rows][C]) and packed pre-activations y (planes [2][rows][2C]) -> dy planes [2][rows][2C] (packed order):
//   z = sigmoid(g) tanh(f);  dg = dz * tanh(f) * sg (1 - sg);  df = dz * sg * (1 - tanh(f)^2)
__global__ void k_gate_bwd(const float* __restrict__ dz, const uint16_t* __restrict__ y, uint16_t* __restrict__ dy,
                           long long rows, int C, int gate_tile, int prec) {
  const int half = gate_tile / 2;
  const long long n4 = rows * C / 4;
  const size_t plane = (size_t)rows * 2 * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i * 4;
    const long long row = e / C;
    const int c = (int)(e % C);
    const int ng = (c / half) * gate_tile + (c % half);
    const size_t off = (size_t)row * 2 * C + ng;
    float g[4], f[4], d4[4], dg[4], df[4];
    fd_load_planes<4>(y, plane, off, g, prec);
    fd_load_planes<4>(y, plane, off + half, f, prec);
    fd_load_f32<4>(dz + e, d4);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float sg = fd_sigmoid(g[k]), th = fd_tanh(f[k]);
      dg[k] = d4[k] * th * sg * (1.f - sg);
      df[k] = d4[k] * sg * (1.f - th * th);
    }
    fd_store_planes<4>(dy, plane, off, dg, prec);
    fd_store_planes<4>(dy, plane, off + half, df, prec);
  }
}

// grad (fp32 [n]) masked by the sign of the forward activation (planes): out planes = split(grad * (act > 0) * scale)
__global__ void k_relu_bwd(const float* __restrict__ grad, const uint16_t* __restrict__ act, uint16_t* __restrict__ out,
                           long long n, float scale, int prec) {
  const long long n4 = n / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i * 4;
    float a[4], g[4];
    fd_load_planes<4>(act, (size_t)n, (size_t)e, a, prec);
    fd_load_f32<4>(grad + e, g);
#pragma unroll
    for (int k = 0; k < 4; ++k) g[k] = a[k] > 0.f ? g[k] * scale : 0.f;
    fd_store_planes<4>(out, (size_t)n, (size_t)e, g, prec);
  }
}

// LeakyReLU backward with an optional addend (autograd of `x + c2(lrelu(c1(lrelu(x))))`, nsf_hifigan/models.py:103-110):
//   v = grad * (act > 0 ? 1 : slope) * scale + addend      act = planes of lrelu(x) (same sign as x)
// -> out_f32 and / or out_planes (either may be null, not both)
__global__ void k_lrelu_bwd(const float* __restrict__ grad, const uint16_t* __restrict__ act,
                            const float* __restrict__ addend, float* __restrict__ out_f32,
                            uint16_t* __restrict__ out_planes, long long n, float slope, float scale, int prec) {
  const long long n4 = n / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i * 4;
    float a[4], g[4];
    fd_load_planes<4>(act, (size_t)n, (size_t)e, a, prec);
    fd_load_f32<4>(grad + e, g);
#pragma unroll
    for (int k = 0; k < 4; ++k) g[k] = (a[k] > 0.f ? g[k] : g[k] * slope) * scale;
    if (addend != nullptr) {
      float r[4];
      fd_load_f32<4>(addend + e, r);
#pragma unroll
      for (int k = 0; k < 4; ++k) g[k] += r[k];
    }
    if (out_f32 != nullptr) fd_store_f32<4>(out_f32 + e, g);
    if (out_planes != nullptr) fd_store_planes<4>(out_planes, (size_t)n, (size_t)e, g, prec);
  }
}

// column sums per batch item: in (planes [2][B][T][N] or fp32 [B][T][N]) -> out[b][n] += scale * sum_t in[b,t,n]
// (out must be zero-initialised; fp32 atomics over the row chunks)
__global__ void k_colsum(const uint16_t* __restrict__ planes, const float* __restrict__ f32, float* __restrict__ out,
                         int B, int T, int N, float scale, int rows_per_block, int prec) {
  const int b = blockIdx.z;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int t0 = blockIdx.y * rows_per_block;
  const int t1 = min(T, t0 + rows_per_block);
  const size_t plane = (size_t)B * T * N;
  float acc = 0.f;
  for (int t = t0; t < t1; ++t) {
    const size_t off = ((size_t)b * T + t) * N + n;
    acc += planes != nullptr ? fd_combine(planes[off], planes[plane + off], prec) : f32[off];
  }
  atomicAdd(out + (size_t)b * N + n, acc * scale);
}

// column sums over the first / last `e` time steps of every item: out[edge][b][n] += scale * sum in[b,t,n],
// edge 0: t in [0,e), edge 1: t in [T-e,T)   (conv-tap edge corrections of the step-vector term in the weight gradient)
__global__ void k_colsum_edges(const uint16_t* __restrict__ planes, float* __restrict__ out, int B, int T, int N, int e,
                               float scale, int rows_per_block, int chunks, int prec) {
  const int b = blockIdx.z;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int edge = blockIdx.y / chunks, chunk = blockIdx.y % chunks;
  const int base = edge == 0 ? 0 : T - e;
  const int t0 = base + chunk * rows_per_block;
  const int t1 = min(base + e, t0 + rows_per_block);
  const size_t plane = (size_t)B * T * N;
  float acc = 0.f;
  for (int t = t0; t < t1; ++t) {
    const size_t off = ((size_t)b * T + t) * N + n;
    acc += fd_combine(planes[off], planes[plane + off], prec);
  }
  atomicAdd(out + ((size_t)edge * B + b) * N + n, acc * scale);
}

// out[i] = scale * sum_b in[b][i]
__global__ void k_reduce_batch(const float* __restrict__ in, float* __restrict__ out, int B, long long n, float scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc += in[(size_t)b * n + i];
    out[i] = acc * scale;
  }
}

inline int grid1d(long long work, int block = 256, int cap = 148 * 16) {
  long long g = (work + block - 1) / block;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" {

int fd_fold_transpose(const uint16_t* src_planes, const float* src_f32, const uint16_t* aux_planes, const float* addvec,
                      int add_bstride, uint16_t* dst, int B, int T, int C, int Tp, int pad, float scale, int mode,
                      int gate_tile, int prec, int dst_rows, int dst_row0, void* stream) {
  FD_DEVICE_GUARD();
  FD_REQUIRE(Tp >= T + pad && pad >= 0, "fd_fold_transpose: Tp=%d too small for T=%d pad=%d", Tp, T, pad);
  if (dst_rows <= 0) { dst_rows = C; dst_row0 = 0; }
  FD_REQUIRE(dst_row0 >= 0 && dst_row0 + C <= dst_rows, "fd_fold_transpose: rows [%d,%d) outside the %d-row destination",
             dst_row0, dst_row0 + C, dst_rows);
  FD_REQUIRE(mode >= 0 && mode <= 2, "fd_fold_transpose: bad mode %d", mode);
  dim3 grid((Tp + 31) / 32, (C + 31) / 32, B), block(32, 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 0)
    k_fold_transpose<0><<<grid, block, 0, st>>>(src_planes, nullptr, nullptr, addvec, add_bstride, dst, B, T, C, Tp, pad,
                                                scale, gate_tile, prec, dst_rows, dst_row0);
  else if (mode == 1)
    k_fold_transpose<1><<<grid, block, 0, st>>>(src_planes, nullptr, nullptr, nullptr, 0, dst, B, T, C, Tp, pad, scale,
                                                gate_tile, prec, dst_rows, dst_row0);
  else
    k_fold_transpose<2><<<grid, block, 0, st>>>(nullptr, src_f32, aux_planes, nullptr, 0, dst, B, T, C, Tp, pad, scale,
                                                gate_tile, prec, dst_rows, dst_row0);
  FD_LAUNCHED();
  return 0;
}

int fd_gate_bwd(const float* dz, const uint16_t* y_planes, uint16_t* dy_planes, long long rows, int C, int gate_tile,
                int prec, void* stream) {
  FD_DEVICE_GUARD();
  FD_REQUIRE(C % 4 == 0 && gate_tile % 8 == 0, "fd_gate_bwd: C=%d gate_tile=%d unsupported", C, gate_tile);
  k_gate_bwd<<<grid1d(rows * C / 4), 256, 0, (cudaStream_t)stream>>>(dz, y_planes, dy_planes, rows, C, gate_tile, prec);
  FD_LAUNCHED();
  return 0;
}

int fd_relu_bwd(const float* grad, const uint16_t* act_planes, uint16_t* out_planes, long long n, float scale, int prec,
                void* stream) {
  FD_DEVICE_GUARD();
  FD_REQUIRE(n % 4 == 0, "fd_relu_bwd: n=%lld must be a multiple of 4", n);
  k_relu_bwd<<<grid1d(n / 4), 256, 0, (cudaStream_t)stream>>>(grad, act_planes, out_planes, n, scale, prec);
  FD_LAUNCHED();
  return 0;
}

int fd_lrelu_bwd(const float* grad, const uint16_t* act_planes, const float* addend, float* out_f32,
                 uint16_t* out_planes, long long n, float slope, float scale, int prec, void* stream) {
  FD_DEVICE_GUARD();
  FD_REQUIRE(n > 0 && n % 4 == 0, "fd_lrelu_bwd: n=%lld must be a positive multiple of 4", n);
  FD_REQUIRE(grad != nullptr && act_planes != nullptr && (out_f32 != nullptr || out_planes != nullptr),
             "fd_lrelu_bwd: grad, act_planes and at least one output are required");
  k_lrelu_bwd<<<grid1d(n / 4), 256, 0, (cudaStream_t)stream>>>(grad, act_planes, addend, out_f32, out_planes, n, slope,
                                                               scale, prec);
  FD_LAUNCHED();
  return 0;
}

int fd_colsum(const uint16_t* planes, const float* f32, float* out, int B, int T, int N, float scale, int prec,
              void* stream) {
  FD_DEVICE_GUARD();
  FD_REQUIRE((planes != nullptr) != (f32 != nullptr), "fd_colsum: exactly one of planes / f32 must be given");
  const int rows_per_block = 128;
  dim3 grid((N + 127) / 128, (T + rows_per_block - 1) / rows_per_block, B);
  k_colsum<<<grid, 128, 0, (cudaStream_t)stream>>>(planes, f32, out, B, T, N, scale, rows_per_block, prec);
  FD_LAUNCHED();
  return 0;
}

int fd_colsum_edges(const uint16_t* planes, float* out, int B, int T, int N, int e, float scale, int prec,
                    void* stream) {
  FD_DEVICE_GUARD();
  FD_REQUIRE(planes != nullptr && out != nullptr && e >= 0 && e <= T, "fd_colsum_edges: bad arguments (e=%d T=%d)", e, T);
  if (e == 0) return 0;
  const int rows_per_block = 64;
  const int chunks = (e + rows_per_block - 1) / rows_per_block;
  dim3 grid((N + 127) / 128, 2 * chunks, B);
  k_colsum_edges<<<grid, 128, 0, (cudaStream_t)stream>>>(planes, out, B, T, N, e, scale, rows_per_block, chunks, prec);
  FD_LAUNCHED();
  return 0;
}

int fd_reduce_batch(const float* in, float* out, int B, long long n, float scale, void* stream) {
  FD_DEVICE_GUARD();
  k_reduce_batch<<<grid1d(n), 256, 0, (cudaStream_t)stream>>>(in, out, B, n, scale);
  FD_LAUNCHED();
  return 0;
}

int fd_wavenet_block_bwd(const fd_wavenet_bwd_desc* d, void* stream) {
  FD_DEVICE_GUARD();
  FD_REQUIRE(d != nullptr, "fd_wavenet_block_bwd: null descriptor");
  const int B = d->B, T = d->T, C = d->C, E = d->E, dil = d->dilation;
  FD_REQUIRE(B > 0 && T > 0 && C % 64 == 0 && E % 64 == 0 && dil > 0, "fd_wavenet_block_bwd: bad shape B=%d T=%d C=%d E=%d", B, T, C, E);
  const float inv_sqrt2 = 0.70710678118654752440f;
  const long long rows = (long long)B * T;
  int rc;
  // ---- dy = gate backward of dz = [dx_next | d_skip] . W2, fused into the GEMM's epilogue together with the column
  //      sums of dy (K offset C selects the skip half of W2^T when there is no residual gradient)
  {
    fd_gemm_desc g;
    memset(&g, 0, sizeof(g));
    g.w = d->w2t; g.n_total = C; g.k_total = 2 * C; g.B = B; g.T = T;
    g.w_inv_scale = d->w2t_inv; g.res_scale = 1.f; g.post_scale = 1.f; g.planes_scale = 1.f;
    g.out_planes = d->dy; g.prec = d->prec; g.backend = d->backend;
    g.gate_y = d->y_planes; g.gate_tile = d->gate_tile; g.gate_dil = dil < T ? dil : T;
    g.gate_cs = d->cs_dy; g.gate_cs_edge = d->cs_edge; g.gate_cs_scale = d->inv_S;
    if (d->dx_next == nullptr) {
      g.src[0] = d->dskip; g.src_C[0] = C; g.num_seg = 1; g.w_kshift = C;
      g.seg_src[0] = 0; g.seg_shift[0] = 0; g.seg_coff[0] = 0; g.seg_klen[0] = C;
    } else {
      g.src[0] = d->dx_next; g.src_C[0] = C; g.src[1] = d->dskip; g.src_C[1] = C; g.num_seg = 2;
      g.seg_src[0] = 0; g.seg_klen[0] = C; g.seg_src[1] = 1; g.seg_klen[1] = C;
    }
    rc = fd_gemm_cl_fwd(&g, stream);
    if (rc) return rc;
  }
  // ---- gw2 = [dx_next ; d_skip]^T . z
  {
    fd_wgrad_desc w;
    memset(&w, 0, sizeof(w));
    w.col_src[0] = d->z_planes; w.col_C[0] = C; w.num_col_seg = 1; w.col_seg_width[0] = C;
    w.B = B; w.T = T; w.splits = d->splits2; w.part = d->part2; w.acc_scale = 1.f; w.prec = d->prec;
    float* out = d->gw2;
    int R = 2 * C;
    if (d->dx_next == nullptr) {
      FD_CHECK_CUDA(cudaMemsetAsync(d->gw2, 0, (size_t)C * C * sizeof(float), (cudaStream_t)stream));
      w.row_src[0] = d->dskip; w.row_C[0] = C; w.num_row_seg = 1; w.row_seg_width[0] = C;
      out = d->gw2 + (size_t)C * C; R = C;
    } else {
      w.row_src[0] = d->dx_next; w.row_C[0] = C; w.row_src[1] = d->dskip; w.row_C[1] = C; w.num_row_seg = 2;
      w.row_seg_src[0] = 0; w.row_seg_width[0] = C; w.row_seg_src[1] = 1; w.row_seg_width[1] = C;
    }
    rc = fd_wgrad_cl(&w, stream);
    if (rc) return rc;
    rc = fd_reduce_batch(d->part2, out, d->splits2, (long long)R * C, d->inv_S, stream);
    if (rc) return rc;
  }
  // ---- gw1 = dy^T . [x(t-d) | x(t) | x(t+d) | cond]
  {
    fd_wgrad_desc w;
    memset(&w, 0, sizeof(w));
    w.row_src[0] = d->dy; w.row_C[0] = 2 * C; w.num_row_seg = 1; w.row_seg_width[0] = 2 * C;
    w.col_src[0] = d->x_planes; w.col_C[0] = C; w.col_src[1] = d->cond_planes; w.col_C[1] = E; w.num_col_seg = 4;
    const int sh[3] = {-dil, 0, dil};
    for (int j = 0; j < 3; ++j) { w.col_seg_src[j] = 0; w.col_seg_shift[j] = sh[j]; w.col_seg_width[j] = C; }
    w.col_seg_src[3] = 1; w.col_seg_width[3] = E;
    w.B = B; w.T = T; w.splits = d->splits1; w.part = d->part1; w.acc_scale = 1.f; w.prec = d->prec;
    rc = fd_wgrad_cl(&w, stream);
    if (rc) return rc;
    rc = fd_reduce_batch(d->part1, d->gw1, d->splits1, (long long)2 * C * (3 * C + E), d->inv_S, stream);
    if (rc) return rc;
  }
  // ---- dx_l = conv^T(dy) + dx_next/sqrt2 (mirrored tap shifts, K = 6C)
  {
    fd_gemm_desc g;
    memset(&g, 0, sizeof(g));
    g.src[0] = d->dy; g.src_C[0] = 2 * C; g.w = d->w1t; g.n_total = C; g.k_total = 6 * C; g.B = B; g.T = T; g.num_seg = 3;
    const int sh[3] = {dil, 0, -dil};
    for (int j = 0; j < 3; ++j) { g.seg_src[j] = 0; g.seg_shift[j] = sh[j]; g.seg_klen[j] = 2 * C; }
    g.res_planes = d->dx_next; g.res_scale = inv_sqrt2; g.w_inv_scale = d->w1t_inv; g.post_scale = 1.f; g.planes_scale = 1.f;
    g.out_planes = d->dx_out; g.out_f32 = d->dx_f32; g.prec = d->prec; g.backend = d->backend;
    rc = fd_gemm_cl_fwd(&g, stream);
    if (rc) return rc;
  }
  if (d->d_cond != nullptr) {
    fd_gemm_desc g;
    memset(&g, 0, sizeof(g));
    g.src[0] = d->dy; g.src_C[0] = 2 * C; g.w = d->wct; g.n_total = E; g.k_total = 2 * C; g.B = B; g.T = T; g.num_seg = 1;
    g.seg_klen[0] = 2 * C;
    g.w_inv_scale = d->wct_inv * d->inv_S; g.res_scale = 1.f; g.post_scale = 1.f; g.planes_scale = 1.f;
    g.out_f32 = d->d_cond; g.out_accum = 1; g.prec = d->prec; g.backend = d->backend;
    rc = fd_gemm_cl_fwd(&g, stream);
    if (rc) return rc;
  }
  return fd_colsum(d->dx_out, nullptr, d->cs_dx, B, T, C, d->inv_S, d->prec & 0xF, stream);
}

}  // extern "C"
