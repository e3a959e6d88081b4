// NSF-HiFiGAN pieces that are not GEMM-shaped: the harmonic source module (exact phase scan),
// the 1-channel source convolutions, conv_post + tanh, and the reflect pad of the mel front end.
// All HBM-bound: the source module reads frame-rate f0 and writes one float per audio sample.
#include <curand_kernel.h>
#include "fd_common.cuh"
#include "fd_host.h"

namespace {

constexpr int SG_THREADS = 256;
constexpr int SG_PER = 8;                       // consecutive samples per thread
constexpr int SG_CHUNK = SG_THREADS * SG_PER;   // samples per block
constexpr int SG_MAXH = 16;

// f0 upsampling: F.interpolate(mode="linear", align_corners=False) (models.py:411-413), fp32 like ATen's
// area_pixel_compute_source_index / guard_index_and_lambda.
__device__ __forceinline__ float f0_upsample(const float* __restrict__ f0b, int T, float scale, long long s) {
  float src = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)s, 0.5f)), 0.5f);
  if (src < 0.f) src = 0.f;
  int i0 = (int)src;
  if (i0 > T - 1) i0 = T - 1;
  const int i1 = i0 + (i0 < T - 1 ? 1 : 0);
  float lam = __fsub_rn(src, (float)i0);
  lam = fminf(fmaxf(lam, 0.f), 1.f);
  const float w0 = __fsub_rn(1.f, lam);
  return __fmaf_rn(w0, f0b[i0], __fmul_rn(lam, f0b[i1]));   // ATen's evaluation order (bit-exact vs torch CPU)
}

// rad = ((f0 * h) / sr) % 1 in fp32 (models.py:208, 268-270), returned as an exact 0.64 fixed-point fraction.
// For the first sample of an item rand_ini is added in fp32 first (models.py:211-215).
__device__ __forceinline__ unsigned long long rad_fixed(float f0, int h, float sr, float rand_ini, bool first) {
  float r = fmodf(__fdiv_rn(__fmul_rn(f0, (float)h), sr), 1.f);
  if (first) r = __fadd_rn(r, rand_ini);
  double d = (double)r;
  d -= floor(d);   // integer parts do not change sin(2*pi*phase)
  return __double2ull_rz(d * 18446744073709551616.0);
}

// pass 1: per-chunk sums of the phase increments, [B][nchunk][H]
__global__ void k_sinegen_sums(const float* __restrict__ f0, const float* __restrict__ rand_ini,
                               unsigned long long* __restrict__ sums, int T, long long S, int hop, int H, float sr,
                               int nchunk) {
  const int b = blockIdx.y, chunk = blockIdx.x;
  const float* f0b = f0 + (size_t)b * T;
  const float scale = __fdiv_rn((float)T, (float)S);
  __shared__ unsigned long long red[SG_THREADS / 32];
  const long long s0 = (long long)chunk * SG_CHUNK + (long long)threadIdx.x * SG_PER;
  float fu[SG_PER];
#pragma unroll
  for (int i = 0; i < SG_PER; ++i) fu[i] = (s0 + i < S) ? f0_upsample(f0b, T, scale, s0 + i) : 0.f;
  for (int h = 1; h <= H; ++h) {
    unsigned long long acc = 0;
    const float ri = rand_ini[(size_t)b * H + (h - 1)];
#pragma unroll
    for (int i = 0; i < SG_PER; ++i)
      if (s0 + i < S) acc += rad_fixed(fu[i], h, sr, ri, s0 + i == 0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long t = 0;
      for (int w = 0; w < SG_THREADS / 32; ++w) t += red[w];
      sums[((size_t)b * nchunk + chunk) * H + (h - 1)] = t;
    }
    __syncthreads();
  }
}

// pass 2: exclusive scan over chunks (in place), one thread per (b,h)
__global__ void k_sinegen_scan(unsigned long long* __restrict__ sums, int B, int nchunk, int H) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H) return;
  const int b = i / H, h = i % H;
  unsigned long long run = 0;
  for (int c = 0; c < nchunk; ++c) {
    const size_t o = ((size_t)b * nchunk + c) * H + h;
    const unsigned long long v = sums[o];
    sums[o] = run;
    run += v;
  }
}

// pass 3: phases, sines, uv, noise, Linear(H->1), tanh
__global__ void k_sinegen_out(const float* __restrict__ f0, const float* __restrict__ rand_ini,
                              const unsigned long long* __restrict__ prefix, const float* __restrict__ lin_w,
                              const float* __restrict__ lin_b, const float* __restrict__ noise,
                              float* __restrict__ har, int T, long long S, int hop, int H, float sr, float sine_amp,
                              float noise_std, int nchunk, unsigned long long seed) {
  const int b = blockIdx.y, chunk = blockIdx.x;
  const float* f0b = f0 + (size_t)b * T;
  const float scale = __fdiv_rn((float)T, (float)S);
  __shared__ unsigned long long wsum[SG_THREADS / 32];
  const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
  const long long s0 = (long long)chunk * SG_CHUNK + (long long)threadIdx.x * SG_PER;
  float fu[SG_PER], lin[SG_PER];
#pragma unroll
  for (int i = 0; i < SG_PER; ++i) {
    fu[i] = (s0 + i < S) ? f0_upsample(f0b, T, scale, s0 + i) : 0.f;
    lin[i] = lin_b[0];
  }
  for (int h = 1; h <= H; ++h) {
    const float ri = rand_ini[(size_t)b * H + (h - 1)];
    unsigned long long ph[SG_PER];
    unsigned long long run = 0;
#pragma unroll
    for (int i = 0; i < SG_PER; ++i) {
      if (s0 + i < S) run += rad_fixed(fu[i], h, sr, ri, s0 + i == 0);
      ph[i] = run;   // inclusive within the thread
    }
    // block-wide exclusive scan of the per-thread totals
    unsigned long long incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    unsigned long long base = prefix[((size_t)b * nchunk + chunk) * H + (h - 1)];
    for (int w = 0; w < warp; ++w) base += wsum[w];
    base += incl - run;
    __syncthreads();
    const float wh = lin_w[h - 1];
#pragma unroll
    for (int i = 0; i < SG_PER; ++i) {
      if (s0 + i >= S) continue;
      const unsigned long long p = base + ph[i];
      const double frac = (double)(p >> 11) * (1.0 / 9007199254740992.0);   // 2^-53
      const float arg = (float)(frac * 6.283185307179586476925286766559);
      const float sine = sinf(arg) * sine_amp;
      const float uv = fu[i] > 0.f ? 1.f : 0.f;
      const float namp = uv * noise_std + (1.f - uv) * sine_amp / 3.f;
      float nz;
      if (noise != nullptr) {
        nz = noise[((size_t)b * S + (s0 + i)) * H + (h - 1)];
      } else {
        curandStatePhilox4_32_10_t st;
        curand_init(seed, (unsigned long long)(((size_t)b * S + (s0 + i)) * SG_MAXH + (h - 1)), 0, &st);
        nz = curand_normal(&st);
      }
      lin[i] = fmaf(wh, sine * uv + namp * nz, lin[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < SG_PER; ++i)
    if (s0 + i < S) har[(size_t)b * S + s0 + i] = tanhf(lin[i]);
}

// 1 -> C strided conv of the excitation; w_t [k][C] (tap-major so lanes read consecutive channels).
// HBM-bound (4 B/element written): a thread owns 4 consecutive channels x SC_R consecutive output rows, so a weight
// float4 (L1-resident, coalesced over the channel groups) feeds 4*SC_R FMAs and every store is a float4; the slice of
// the excitation a block needs sits in shared memory.
constexpr int SC_R = 4;
__global__ void k_source_conv(const float* __restrict__ har, const float* __restrict__ w_t,
                              const float* __restrict__ bias, float* __restrict__ out, long long S, long long S_out,
                              int C, int k, int s, int p, int rows_per_block) {
  extern __shared__ float win[];
  const int b = blockIdx.y;
  const int cgroups = C / 4;
  const int cg = threadIdx.x % cgroups, rg = threadIdx.x / cgroups;
  const int wlen = (rows_per_block - 1) * s + k;
  const float4 b4 = *reinterpret_cast<const float4*>(bias + 4 * cg);
  for (long long q0 = (long long)blockIdx.x * rows_per_block; q0 < S_out; q0 += (long long)gridDim.x * rows_per_block) {
    const long long base = q0 * s - p;
    __syncthreads();
    for (int i = threadIdx.x; i < wlen; i += blockDim.x) {
      const long long idx = base + i;
      win[i] = (idx >= 0 && idx < S) ? har[(size_t)b * S + idx] : 0.f;
    }
    __syncthreads();
    const int r0 = rg * SC_R;
    float4 acc[SC_R];
#pragma unroll
    for (int r = 0; r < SC_R; ++r) acc[r] = b4;
    for (int j = 0; j < k; ++j) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(w_t + (size_t)j * C + 4 * cg));
#pragma unroll
      for (int r = 0; r < SC_R; ++r) {
        const float h = win[(r0 + r) * s + j];
        acc[r].x = fmaf(h, w4.x, acc[r].x); acc[r].y = fmaf(h, w4.y, acc[r].y);
        acc[r].z = fmaf(h, w4.z, acc[r].z); acc[r].w = fmaf(h, w4.w, acc[r].w);
      }
    }
#pragma unroll
    for (int r = 0; r < SC_R; ++r)
      if (q0 + r0 + r < S_out)
        *reinterpret_cast<float4*>(out + ((size_t)b * S_out + q0 + r0 + r) * C + 4 * cg) = acc[r];
  }
}

// conv_post (C -> 1, k taps, zero padding k/2) + tanh over split planes; w [k][C] in shared memory
__global__ void k_conv_post(const uint16_t* __restrict__ in_planes, const float* __restrict__ w,
                            const float* __restrict__ bias, float* __restrict__ wav, int B, long long S, int C, int k,
                            int prec) {
  extern __shared__ float ws[];
  for (int i = threadIdx.x; i < k * C; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const size_t plane = (size_t)B * S * C;
  const int half = k / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)B * S;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / S, s = i % S;
    float acc = bias[0];
    for (int j = 0; j < k; ++j) {
      const long long t = s + j - half;
      if (t < 0 || t >= S) continue;
      const size_t off = ((size_t)b * S + t) * C;
      for (int c = 0; c < C; c += 8) {
        float v[8];
        fd_load_planes<8>(in_planes, plane, off + c, v, prec);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc = fmaf(v[e], ws[j * C + c + e], acc);
      }
    }
    wav[i] = tanhf(acc);
  }
}

__global__ void k_reflect_pad_split(const float* __restrict__ wav, uint16_t* __restrict__ planes, int B, long long N,
                                    long long Np, long long pitch, int pad, int prec) {
  const size_t plane = (size_t)B * pitch;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)B * pitch;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / pitch, j = i % pitch;
    float v = 0.f;
    if (j < Np) {
      long long src = j - pad;
      if (src < 0) src = -src;
      if (src >= N) src = 2 * (N - 1) - src;
      v = wav[(size_t)b * N + src];
    }
    uint16_t hi, lo;
    fd_split(v, prec, hi, lo);
    planes[i] = hi;
    planes[plane + i] = lo;
  }
}

}  // namespace

extern "C" {

size_t fd_sinegen_ws_bytes(int B, long long S) {
  const long long nchunk = (S + SG_CHUNK - 1) / SG_CHUNK;
  return (size_t)B * nchunk * SG_MAXH * sizeof(unsigned long long);
}

int fd_sinegen_fwd(const float* f0, const float* lin_w, const float* lin_b, const float* rand_ini,
                   const float* noise, float* har, void* ws, int B, int T, int hop, int H, float sampling_rate,
                   float sine_amp, float noise_std, unsigned long long seed, void* stream) {
  FD_DEVICE_GUARD();
  FD_REQUIRE(H >= 1 && H <= SG_MAXH, "fd_sinegen_fwd: H=%d out of range", H);
  FD_REQUIRE(B > 0 && T > 0 && hop > 0, "fd_sinegen_fwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const long long S = (long long)T * hop;
  const int nchunk = (int)((S + SG_CHUNK - 1) / SG_CHUNK);
  unsigned long long* sums = reinterpret_cast<unsigned long long*>(ws);
  dim3 grid(nchunk, B);
  k_sinegen_sums<<<grid, SG_THREADS, 0, st>>>(f0, rand_ini, sums, T, S, hop, H, sampling_rate, nchunk);
  FD_LAUNCHED();
  k_sinegen_scan<<<(B * H + 127) / 128, 128, 0, st>>>(sums, B, nchunk, H);
  FD_LAUNCHED();
  k_sinegen_out<<<grid, SG_THREADS, 0, st>>>(f0, rand_ini, sums, lin_w, lin_b, noise, har, T, S, hop, H,
                                              sampling_rate, sine_amp, noise_std, nchunk, seed);
  FD_LAUNCHED();
  return 0;
}

int fd_source_conv_fwd(const float* har, const float* w, const float* bias, float* out, int B, long long S, int C,
                       int k, int s, int p, void* stream) {
  FD_DEVICE_GUARD();
  const long long S_out = (S + 2LL * p - k) / s + 1;
  FD_REQUIRE(S_out > 0, "fd_source_conv_fwd: empty output");
  FD_REQUIRE(C % 4 == 0 && C <= 1024 && 1024 % C == 0, "fd_source_conv_fwd: C=%d must divide 1024", C);
  const int rows_per_block = (256 / (C / 4)) * SC_R;
  const int wlen = (rows_per_block - 1) * s + k;
  long long gx = (S_out + rows_per_block - 1) / rows_per_block;
  const long long cap = (148LL * 8 + B - 1) / B;
  if (gx > cap) gx = cap;
  dim3 grid((unsigned)gx, B);
  k_source_conv<<<grid, 256, wlen * sizeof(float), (cudaStream_t)stream>>>(har, w, bias, out, S, S_out, C, k, s, p,
                                                                           rows_per_block);
  FD_LAUNCHED();
  return 0;
}

int fd_conv_post_fwd(const uint16_t* in_planes, const float* w, const float* bias, float* wav, int B, long long S,
                     int C, int k, int prec, void* stream) {
  FD_DEVICE_GUARD();
  FD_REQUIRE(C % 8 == 0, "fd_conv_post_fwd: C=%d must be a multiple of 8", C);
  const long long n = (long long)B * S;
  long long g = (n + 255) / 256;
  if (g > 148 * 32) g = 148 * 32;
  k_conv_post<<<(unsigned)g, 256, k * C * sizeof(float), (cudaStream_t)stream>>>(in_planes, w, bias, wav, B, S, C, k,
                                                                                 prec);
  FD_LAUNCHED();
  return 0;
}

int fd_reflect_pad_split(const float* wav, uint16_t* planes, int B, long long N, int pad, int prec, void* stream) {
  FD_DEVICE_GUARD();
  FD_REQUIRE(pad < N, "fd_reflect_pad_split: pad=%d must be smaller than N=%lld", pad, N);
  const long long Np = N + 2LL * pad;
  const long long pitch = (Np + 7) / 8 * 8;
  long long g = ((long long)B * pitch + 255) / 256;
  if (g > 148 * 32) g = 148 * 32;
  k_reflect_pad_split<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(wav, planes, B, N, Np, pitch, pad, prec);
  FD_LAUNCHED();
  return 0;
}

}  // extern "C"
