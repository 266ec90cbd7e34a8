// Kernels of HQAViT's lateral CNN path and SplitFusion (scope row f-1; H = HQAViT_CIFAR100.py):
//   - BatchNorm2d (train: batch statistics + running-stat update, eval: running statistics) fused with GELU  (H:753-775)
//   - 3x3 stride-2 convolutions as im2col + GEMM (im2col / col2im / weight packing)                            (H:752, 759)
//   - LayerNorm fused with its consumer: GELU (LMFAdapter tail, H:833-834) or beta * LN + residual (RRCV, H:900-903)
//   - SplitFusion's row-wise parts as two kernels per direction around the two GEMMs                            (H:945-963)
// Every kernel is HBM-bound: rows are walked once, 16 B accesses, fp32 math, statistics in fp32.
#include "kernels.h"

namespace {

// ------------------------------------------------------------------------------------------------ small helpers
template <typename T> struct Vec8;
template <> struct Vec8<float> {
  static __device__ __forceinline__ void ld(const float* p, float* v) { load_vec<8>(p, v); }
  static __device__ __forceinline__ void st(float* p, const float* v) { store_vec<8>(p, v); }
};
template <> struct Vec8<bf16> {
  static __device__ __forceinline__ void ld(const bf16* p, float* v) { load_vec<8>(p, v); }
  static __device__ __forceinline__ void st(bf16* p, const float* v) { store_vec<8>(p, v); }
};

// keep-mask scale for 8 consecutive elements starting at element index e0 (multiple of 8): the scheme of drop_rows / the GEMM
// epilogues (drop_keep8: ONE Philox call per 8 elements, 16 random bits each; was two calls with 24 bits) -- so the mask of
// SplitFusion's Dropout is Site(seed, offset, site, p).keep_rows(rows, C) in tests/dropout_masks.py
__device__ __forceinline__ void dropout_scale8(const unsigned long long* rng, uint32_t site, unsigned long long e0, float p,
                                               float* sc) {
  DropP d;
  d.p = p; d.rng = rng; d.site = site;
  const DropState st = drop_state(d);
  drop_keep8(st, e0 >> 3, sc);
}

__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

// ------------------------------------------------------------------------------------------------ BatchNorm
// sums[0..C) += sum_rows (x - shift), sums[C..2C) += sum_rows (x - shift)^2, shift = x[0, c]  (guards the variance
// against cancellation when |mean| >> std).  Thread = 8 channels x a strided set of rows.
// The per-thread fp32 partials are accumulated as 64-bit FIXED-POINT integers (quantum 2^-24): integer addition is associative, so
// the batch statistics -- and with them the whole forward pass -- are bit-identical from run to run whatever order the atomics land
// in (fp32 atomics here made HQAViT's bf16 logits differ by 3e-3 between identical runs: the TokenLearner gate amplifies ulps).
constexpr float BN_FIX = 16777216.f;   // 2^24
__device__ __forceinline__ unsigned long long bn_to_fix(float v) { return (unsigned long long)__float2ll_rn(v * BN_FIX); }
__device__ __forceinline__ float bn_from_fix(unsigned long long v) { return (float)((double)(long long)v * (1.0 / 16777216.0)); }
template <typename T>
__global__ void __launch_bounds__(256) bn_stats_kernel(const T* __restrict__ x, long rows, int C, unsigned long long* __restrict__ sums) {
  QV_PDL_ENTRY();
  extern __shared__ unsigned long long red_fix[];   // [2C]
  const int lpr = C / 8, cv = (threadIdx.x % lpr) * 8, r0 = threadIdx.x / lpr, rpp = 256 / lpr;
  for (int i = threadIdx.x; i < 2 * C; i += 256) red_fix[i] = 0ull;
  float sh[8], s[8], q[8];
  Vec8<T>::ld(x + cv, sh);
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  for (long row = (long)blockIdx.x * rpp + r0; row < rows; row += (long)gridDim.x * rpp) {
    float v[8];
    Vec8<T>::ld(x + row * C + cv, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float d = v[i] - sh[i]; s[i] += d; q[i] = fmaf(d, d, q[i]); }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) { atomicAdd(red_fix + cv + i, bn_to_fix(s[i])); atomicAdd(red_fix + C + cv + i, bn_to_fix(q[i])); }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += 256) atomicAdd(sums + i, red_fix[i]);
}

// mr[0..C) = mean, mr[C..2C) = rstd; train: from the batch sums (+ running-stat update, H: nn.BatchNorm2d momentum 0.1,
// unbiased running variance), eval: from the running statistics.
template <typename T>
__global__ void bn_finalize_kernel(const T* __restrict__ x, const float* __restrict__ sums, long rows, int C, float eps,
                                   float momentum, int train, float* running_mean, float* running_var,
                                   long long* num_batches, float* __restrict__ mr) {
  QV_PDL_ENTRY();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (train) {
    const unsigned long long* fix = reinterpret_cast<const unsigned long long*>(sums);   // written by bn_stats_kernel
    const float n = (float)rows;
    const float d = bn_from_fix(fix[c]) / n;
    const float mean = ldf(x + c) + d;
    const float var = fmaxf(bn_from_fix(fix[C + c]) / n - d * d, 0.f);
    mr[c] = mean;
    mr[C + c] = rsqrtf(var + eps);
    if (running_mean) {
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * (n / fmaxf(n - 1.f, 1.f));
      if (c == 0 && num_batches) *num_batches += 1;
    }
  } else {
    mr[c] = running_mean[c];
    mr[C + c] = rsqrtf(running_var[c] + eps);
  }
}

template <typename T, bool GELU>
__global__ void __launch_bounds__(256) bn_apply_kernel(const T* __restrict__ x, long nvec, int C, const float* __restrict__ mr,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       T* __restrict__ y) {
  QV_PDL_ENTRY();
  const int lpr = C / 8;
  if (256 % lpr == 0) {
    // the thread stride (a multiple of 256 vectors) is a multiple of the vectors per row: a thread stays on the same 8
    // channels, so the per-channel constants are loaded once and the body is 8 FMAs (the 64-bit modulo + 32 parameter
    // loads per vector of the generic loop below cost more than the normalisation itself)
    const int cv = (threadIdx.x % lpr) * 8;
    float mu[8], sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { mu[j] = mr[cv + j]; sc[j] = mr[C + cv + j] * gamma[cv + j]; sh[j] = beta[cv + j]; }
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < nvec; i += (long)gridDim.x * 256) {
      float v[8];
      Vec8<T>::ld(x + i * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float o = fmaf(v[j] - mu[j], sc[j], sh[j]);      // centre first: no cancellation when |mean| >> std
        v[j] = GELU ? gelu_t<T>(o) : o;
      }
      Vec8<T>::st(y + i * 8, v);
    }
    return;
  }
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < nvec; i += (long)gridDim.x * 256) {
    const int cv = (int)(i % lpr) * 8;
    float v[8];
    Vec8<T>::ld(x + i * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float o = (v[j] - mr[cv + j]) * mr[C + cv + j] * gamma[cv + j] + beta[cv + j];
      v[j] = GELU ? gelu_t<T>(o) : o;
    }
    Vec8<T>::st(y + i * 8, v);
  }
}

// sums[0..C) += sum g * xhat (dgamma), sums[C..2C) += sum g (dbeta);  g = dy * gelu'(bn(x)) when GELU
template <typename T, bool GELU>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const T* __restrict__ x, const T* __restrict__ dy, long rows, int C,
                                                            const float* __restrict__ mr, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ sums) {
  QV_PDL_ENTRY();
  extern __shared__ float red[];   // [2C]
  const int lpr = C / 8, cv = (threadIdx.x % lpr) * 8, r0 = threadIdx.x / lpr, rpp = 256 / lpr;
  for (int i = threadIdx.x; i < 2 * C; i += 256) red[i] = 0.f;
  float mean[8], rstd[8], gm[8], bt[8], a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mean[i] = mr[cv + i]; rstd[i] = mr[C + cv + i]; gm[i] = gamma[cv + i]; bt[i] = beta[cv + i];
    a[i] = b[i] = 0.f;
  }
  for (long row = (long)blockIdx.x * rpp + r0; row < rows; row += (long)gridDim.x * rpp) {
    float v[8], g[8];
    Vec8<T>::ld(x + row * C + cv, v);
    Vec8<T>::ld(dy + row * C + cv, g);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xh = (v[i] - mean[i]) * rstd[i];
      if (GELU) g[i] *= gelu_grad_t<T>(fmaf(xh, gm[i], bt[i]));
      a[i] = fmaf(g[i], xh, a[i]);
      b[i] += g[i];
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) { atomicAdd(red + cv + i, a[i]); atomicAdd(red + C + cv + i, b[i]); }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += 256) atomicAdd(sums + i, red[i]);
}

template <typename T, bool GELU>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const T* __restrict__ x, const T* __restrict__ dy, long nvec, long rows,
                                                           int C, const float* __restrict__ mr, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, const float* __restrict__ sums,
                                                           int train, T* __restrict__ dx, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta) {
  QV_PDL_ENTRY();
  const int lpr = C / 8;
  const float invn = train ? 1.f / (float)rows : 0.f;
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += 256) { dgamma[c] += sums[c]; dbeta[c] += sums[C + c]; }
  }
  if (256 % lpr == 0) {       // see bn_apply_kernel: the thread keeps its 8 channels, constants live in registers
    const int cv = (threadIdx.x % lpr) * 8;
    float mu[8], rs[8], gm[8], bt[8], s0[8], s1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mu[j] = mr[cv + j]; rs[j] = mr[C + cv + j]; gm[j] = gamma[cv + j]; bt[j] = beta[cv + j];
      s0[j] = invn * sums[cv + j]; s1[j] = invn * sums[C + cv + j];
    }
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < nvec; i += (long)gridDim.x * 256) {
      float v[8], g[8];
      Vec8<T>::ld(x + i * 8, v);
      Vec8<T>::ld(dy + i * 8, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (v[j] - mu[j]) * rs[j];
        if (GELU) g[j] *= gelu_grad_t<T>(fmaf(xh, gm[j], bt[j]));
        v[j] = gm[j] * rs[j] * (g[j] - (s1[j] + xh * s0[j]));
      }
      Vec8<T>::st(dx + i * 8, v);
    }
    return;
  }
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < nvec; i += (long)gridDim.x * 256) {
    const int cv = (int)(i % lpr) * 8;
    float v[8], g[8];
    Vec8<T>::ld(x + i * 8, v);
    Vec8<T>::ld(dy + i * 8, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float rs = mr[C + cv + j], gm = gamma[cv + j];
      const float xh = (v[j] - mr[cv + j]) * rs;
      if (GELU) g[j] *= gelu_grad_t<T>(fmaf(xh, gm, beta[cv + j]));
      v[j] = gm * rs * (g[j] - invn * (sums[C + cv + j] + xh * sums[cv + j]));
    }
    Vec8<T>::st(dx + i * 8, v);
  }
}

// ------------------------------------------------------------------------------------------------ 3x3 stride-2 pad-1 convs
// image [B, Cin, S, S] fp32 (NCHW) -> col [B * Ho * Wo, Kp], k = (ky * 3 + kx) * Cin + cin, zero padded to Kp.
// One thread per output pixel (adjacent threads = adjacent pixels of a row: neighbouring 3 x 3 windows share cache
// lines); the Kp-wide row is written as 8-element vectors.  KP8 = Kp / 8 (4 for the 3-channel stem).
template <typename T, int KP8, int CIN>
__global__ void __launch_bounds__(256) im2col_img_kernel(const float* __restrict__ img, int B, int Cin_rt, int S, T* __restrict__ col) {
  QV_PDL_ENTRY();
  const int Cin = CIN ? CIN : Cin_rt;     // compile-time channel count keeps the row in registers
  const int Ho = S / 2;
  const long rows = (long)B * Ho * Ho;
  for (long row = (long)blockIdx.x * 256 + threadIdx.x; row < rows; row += (long)gridDim.x * 256) {
    const int ox = (int)(row % Ho), oy = (int)((row / Ho) % Ho);
    const long b = row / ((long)Ho * Ho);
    float v[KP8 * 8];
#pragma unroll
    for (int k = 0; k < KP8 * 8; ++k) v[k] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int y = 2 * oy - 1 + t / 3, x = 2 * ox - 1 + t % 3;
      if (y >= 0 && y < S && x >= 0 && x < S) {
#pragma unroll
        for (int cin = 0; cin < (CIN ? CIN : 3); ++cin)
          if (cin < Cin && t * Cin + cin < KP8 * 8) v[t * Cin + cin] = img[((b * Cin + cin) * S + y) * S + x];
      }
    }
#pragma unroll
    for (int q = 0; q < KP8; ++q) Vec8<T>::st(col + row * (KP8 * 8) + q * 8, v + q * 8);
  }
}
// x [B, Hi, Hi, Cin] (NHWC, Cin % 8 == 0) -> col [B * Ho * Ho, 9 * Cin]
template <typename T>
__global__ void __launch_bounds__(256) im2col_nhwc_kernel(const T* __restrict__ x, int B, int Hi, int Cin, T* __restrict__ col) {
  QV_PDL_ENTRY();
  const int Ho = Hi / 2, cv = Cin / 8;
  const long total = (long)B * Ho * Ho * 9 * cv;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
    const int c8 = (int)(i % cv) * 8;
    const int t = (int)((i / cv) % 9), ky = t / 3, kx = t % 3;
    const long row = i / (9 * cv);
    const int ox = (int)(row % Ho), oy = (int)((row / Ho) % Ho);
    const long b = row / ((long)Ho * Ho);
    const int y = 2 * oy - 1 + ky, xx = 2 * ox - 1 + kx;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (y >= 0 && y < Hi && xx >= 0 && xx < Hi) Vec8<T>::ld(x + ((b * Hi + y) * Hi + xx) * Cin + c8, v);
    Vec8<T>::st(col + row * 9 * Cin + t * Cin + c8, v);
  }
}
// dx [B, Hi, Hi, Cin] = gather of dcol [B * Ho * Ho, 9 * Cin] (each input pixel feeds <= 2 x 2 output taps)
template <typename T>
__global__ void __launch_bounds__(256) col2im_nhwc_kernel(const T* __restrict__ dcol, int B, int Hi, int Cin, T* __restrict__ dx) {
  QV_PDL_ENTRY();
  const int Ho = Hi / 2, cv = Cin / 8;
  const long total = (long)B * Hi * Hi * cv;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
    const int c8 = (int)(i % cv) * 8;
    const long pix = i / cv;
    const int xx = (int)(pix % Hi), y = (int)((pix / Hi) % Hi);
    const long b = pix / ((long)Hi * Hi);
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int ty = y + 1 - ky;
      if (ty < 0 || (ty & 1) || ty / 2 >= Ho) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int tx = xx + 1 - kx;
        if (tx < 0 || (tx & 1) || tx / 2 >= Ho) continue;
        float v[8];
        Vec8<T>::ld(dcol + ((b * Ho + ty / 2) * Ho + tx / 2) * 9 * Cin + (ky * 3 + kx) * Cin + c8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] += v[j];
      }
    }
    Vec8<T>::st(dx + pix * Cin + c8, a);
  }
}
// W [N, Cin, 3, 3] -> Wp [N, Kp] with k = (ky * 3 + kx) * Cin + cin (zero padded); unpack: dW += dWp
__global__ void conv_w_pack_kernel(const float* __restrict__ W, int N, int Cin, int Kp, float* __restrict__ Wp) {
  QV_PDL_ENTRY();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * Kp) return;
  const int n = i / Kp, k = i % Kp;
  float v = 0.f;
  if (k < 9 * Cin) { const int cin = k % Cin, t = k / Cin; v = W[(n * Cin + cin) * 9 + t]; }
  Wp[i] = v;
}
__global__ void conv_w_unpack_add_kernel(const float* __restrict__ dWp, int N, int Cin, int Kp, float* __restrict__ dW) {
  QV_PDL_ENTRY();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * Cin * 9) return;
  const int t = i % 9, cin = (i / 9) % Cin, n = i / (9 * Cin);
  dW[i] += dWp[n * Kp + t * Cin + cin];
}

// ------------------------------------------------------------------------------------------------ LayerNorm + consumer
// y = [resid +] [*scale *] f(LN(x)),  f = GELU when gelu_out.  One warp per row.
template <typename T, typename TY, int EPL>
__global__ void __launch_bounds__(256) rowln_fwd_kernel(const T* __restrict__ x, long rows, int C, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps, int gelu_out,
                                                        const TY* __restrict__ resid, const float* __restrict__ scale,
                                                        TY* __restrict__ y, float* __restrict__ y32, float* __restrict__ stats) {
  QV_PDL_ENTRY();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, c0 = lane * EPL;
  const bool act = c0 < C;
  const float invC = 1.f / (float)C;
  const float sc = scale ? *scale : 1.f;
  float gm[EPL], bt[EPL];
#pragma unroll
  for (int i = 0; i < EPL; ++i) gm[i] = bt[i] = 0.f;
  if (act) { load_vec<EPL>(gamma + c0, gm); load_vec<EPL>(beta + c0, bt); }
  for (long row = (long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long)gridDim.x * wpb) {
    float v[EPL], r[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) v[i] = r[i] = 0.f;
    if (act) {
      load_vec<EPL>(x + row * C + c0, v);
      if (resid) load_vec<EPL>(resid + row * C + c0, r);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) s += v[i];
    const float mean = warp_sum(s) * invC;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) { const float d = act ? v[i] - mean : 0.f; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) * invC + eps);
    if (lane == 0) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
      float o = (v[i] - mean) * rstd * gm[i] + bt[i];
      if (gelu_out) o = gelu_t<T>(o);
      v[i] = r[i] + sc * o;
    }
    if (act) {
      store_vec<EPL>(y + row * C + c0, v);
      if (y32) store_vec<EPL>(y32 + row * C + c0, v);
    }
  }
}

// dx = LN-backward(scale * dy * f'(LN(x)));  dgamma / dbeta accumulated;  dscale += sum dy * f(LN(x)) when scale.
template <typename T, typename TD, int EPL>
__global__ void __launch_bounds__(256, 3) rowln_bwd_kernel(const T* __restrict__ x, const TD* __restrict__ dy, long rows, int C,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        const float* __restrict__ stats, int gelu_out,
                                                        const float* __restrict__ scale, float* __restrict__ dscale,
                                                        T* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  QV_PDL_ENTRY();
  __shared__ float red[2][8][256];
  __shared__ float sred[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5, c0 = lane * EPL;
  const bool act = c0 < C;
  const float invC = 1.f / (float)C;
  const float sc = scale ? *scale : 1.f;
  float gm[EPL], bt[EPL], ag[EPL], ab[EPL], asc = 0.f;
#pragma unroll
  for (int i = 0; i < EPL; ++i) gm[i] = bt[i] = ag[i] = ab[i] = 0.f;
  if (act) { load_vec<EPL>(gamma + c0, gm); load_vec<EPL>(beta + c0, bt); }
  for (long row = (long)blockIdx.x * wpb + warp; row < rows; row += (long)gridDim.x * wpb) {
    float xv[EPL], dv[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) xv[i] = dv[i] = 0.f;
    if (act) { load_vec<EPL>(x + row * C + c0, xv); load_vec<EPL>(dy + row * C + c0, dv); }
    const float mean = stats[2 * row], rstd = stats[2 * row + 1];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
      xv[i] = act ? (xv[i] - mean) * rstd : 0.f;        // xhat
      const float o = fmaf(xv[i], gm[i], bt[i]);
      if (scale) asc += dv[i] * (gelu_out ? gelu_t<T>(o) : o);
      float d = sc * dv[i];
      if (gelu_out) d *= gelu_grad_t<T>(o);
      ag[i] += d * xv[i];
      ab[i] += d;
      dv[i] = d * gm[i];
      s1 += dv[i];
      s2 += dv[i] * xv[i];
    }
    const float m1 = warp_sum(s1) * invC, m2 = warp_sum(s2) * invC;
#pragma unroll
    for (int i = 0; i < EPL; ++i) dv[i] = rstd * (dv[i] - m1 - xv[i] * m2);
    if (act) store_vec<EPL>(dx + row * C + c0, dv);
  }
#pragma unroll
  for (int i = 0; i < EPL; ++i) { red[0][warp][lane * EPL + i] = ag[i]; red[1][warp][lane * EPL + i] = ab[i]; }
  asc = warp_sum(asc);
  if (lane == 0) sred[warp] = asc;
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < wpb; ++w) { a += red[0][w][c]; b += red[1][w][c]; }
    atomicAdd(dgamma + c, a);
    atomicAdd(dbeta + c, b);
  }
  if (dscale && threadIdx.x == 0) {
    float a = 0.f;
    for (int w = 0; w < wpb; ++w) a += sred[w];
    atomicAdd(dscale, a);
  }
}

// ------------------------------------------------------------------------------------------------ SplitFusion
// K1:  s = T + R;  g_in = LN_gate(s) -> T-type;  cat = [T | R] -> T-type [rows, 2C];  gate stats kept.
template <typename T, int EPL>
__global__ void __launch_bounds__(256) sf_pre_fwd_kernel(const float* __restrict__ Tin, const float* __restrict__ R, long rows, int C,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         T* __restrict__ g_in, T* __restrict__ cat, float* __restrict__ stats) {
  QV_PDL_ENTRY();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, c0 = lane * EPL;
  const bool act = c0 < C;
  const float invC = 1.f / (float)C;
  float gm[EPL], bt[EPL];
#pragma unroll
  for (int i = 0; i < EPL; ++i) gm[i] = bt[i] = 0.f;
  if (act) { load_vec<EPL>(gamma + c0, gm); load_vec<EPL>(beta + c0, bt); }
  for (long row = (long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long)gridDim.x * wpb) {
    float t[EPL], r[EPL], v[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) t[i] = r[i] = 0.f;
    if (act) { load_vec<EPL>(Tin + row * C + c0, t); load_vec<EPL>(R + row * C + c0, r); }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) { v[i] = t[i] + r[i]; s += v[i]; }
    const float mean = warp_sum(s) * invC;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) { const float d = act ? v[i] - mean : 0.f; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) * invC + 1e-5f);
    if (lane == 0) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
#pragma unroll
    for (int i = 0; i < EPL; ++i) v[i] = (v[i] - mean) * rstd * gm[i] + bt[i];
    if (act) {
      store_vec<EPL>(g_in + row * C + c0, v);
      store_vec<EPL>(cat + row * 2 * C + c0, t);
      store_vec<EPL>(cat + row * 2 * C + C + c0, r);
    }
  }
}

struct SfP {
  const float* Tin; const float* R; const void* glin; const void* cpre;
  const float *cat_g, *cat_b, *fin_g, *fin_b, *fw;   // cat_mlp.1, final_norm, fusion_weights[2]
  float drop_p; const unsigned long long* rng; uint32_t site;
  long rows; int C;
};

template <typename T, int EPL>
struct SfRow {
  float t[EPL], r[EPL], gate[EPL], xc[EPL], lnc[EPL], m[EPL], dsc[EPL], u[EPL];
  float rstd_c;
  // recomputes everything between the two GEMM outputs and u for one row (shared by forward and backward)
  __device__ __forceinline__ void compute(const SfP& p, long row, int c0, bool act, const float* cg, const float* cb, float w0,
                                          float w1, float mean_c_in, float rstd_c_in, bool have_stats, float* mean_c_out) {
    const float invC = 1.f / (float)p.C;
    float gl[EPL], cp[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) t[i] = r[i] = gl[i] = cp[i] = 0.f;
    if (act) {
      load_vec<EPL>(p.Tin + row * p.C + c0, t);
      load_vec<EPL>(p.R + row * p.C + c0, r);
      load_vec<EPL>(static_cast<const T*>(p.glin) + row * p.C + c0, gl);
      load_vec<EPL>(static_cast<const T*>(p.cpre) + row * p.C + c0, cp);
    }
    float mean_c = mean_c_in;
    rstd_c = rstd_c_in;
    if (!have_stats) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < EPL; ++i) s += cp[i];
      mean_c = warp_sum(s) * invC;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < EPL; ++i) { const float d = act ? cp[i] - mean_c : 0.f; q += d * d; }
      rstd_c = rsqrtf(warp_sum(q) * invC + 1e-5f);
    }
    *mean_c_out = mean_c;
#pragma unroll
    for (int i = 0; i < EPL; ++i) dsc[i] = 1.f;
    if (p.drop_p > 0.f && act) {
      if (EPL == 8) dropout_scale8(p.rng, p.site, (unsigned long long)row * p.C + c0, p.drop_p, dsc);
      else {   // 4-wide lanes: take the half of the 8-group this lane owns
        float tmp[8];
        const unsigned long long e = (unsigned long long)row * p.C + c0;
        dropout_scale8(p.rng, p.site, e & ~7ull, p.drop_p, tmp);
#pragma unroll
        for (int i = 0; i < EPL; ++i) dsc[i] = tmp[(e & 4) + i];
      }
    }
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
      xc[i] = act ? (cp[i] - mean_c) * rstd_c : 0.f;
      lnc[i] = fmaf(xc[i], cg[i], cb[i]);
      m[i] = gelu_t<T>(lnc[i]) * dsc[i];
      gate[i] = sigmoid_f(gl[i]);
      u[i] = act ? w0 * (t[i] + gate[i] * r[i]) + w1 * (t[i] + m[i]) : 0.f;
    }
  }
};

__device__ __forceinline__ void softmax2(const float* fw, float* w0, float* w1) {
  const float a = fw[0], b = fw[1], mx = fmaxf(a, b);
  const float ea = expf(a - mx), eb = expf(b - mx), inv = 1.f / (ea + eb);
  *w0 = ea * inv; *w1 = eb * inv;
}

// K2: out = LN_final(w0 * (T + sigmoid(glin) * R) + w1 * (T + dropout(gelu(LN_cat(cpre)))))
#ifndef SFF_M
#define SFF_M 3
#endif
#ifndef SFP_M
#define SFP_M 3
#endif
template <typename T, int EPL>
__global__ void __launch_bounds__(256, SFF_M) sf_post_fwd_kernel(SfP p, float* __restrict__ out, float* __restrict__ cat_stats,
                                                          float* __restrict__ fin_stats) {
  QV_PDL_ENTRY();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, c0 = lane * EPL;
  const bool act = c0 < p.C;
  const float invC = 1.f / (float)p.C;
  float w0, w1;
  softmax2(p.fw, &w0, &w1);
  float cg[EPL], cb[EPL], fg[EPL], fb[EPL];
#pragma unroll
  for (int i = 0; i < EPL; ++i) cg[i] = cb[i] = fg[i] = fb[i] = 0.f;
  if (act) {
    load_vec<EPL>(p.cat_g + c0, cg); load_vec<EPL>(p.cat_b + c0, cb);
    load_vec<EPL>(p.fin_g + c0, fg); load_vec<EPL>(p.fin_b + c0, fb);
  }
  SfRow<T, EPL> R;
  for (long row = (long)blockIdx.x * wpb + (threadIdx.x >> 5); row < p.rows; row += (long)gridDim.x * wpb) {
    float mean_c;
    R.compute(p, row, c0, act, cg, cb, w0, w1, 0.f, 0.f, false, &mean_c);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) s += R.u[i];
    const float mean = warp_sum(s) * invC;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) { const float d = act ? R.u[i] - mean : 0.f; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) * invC + 1e-5f);
    if (lane == 0) {
      cat_stats[2 * row] = mean_c; cat_stats[2 * row + 1] = R.rstd_c;
      fin_stats[2 * row] = mean; fin_stats[2 * row + 1] = rstd;
    }
    float o[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) o[i] = (R.u[i] - mean) * rstd * fg[i] + fb[i];
    if (act) store_vec<EPL>(out + row * p.C + c0, o);
  }
}

// K2 backward: dT_part (fp32), dR_part, dglin, dcpre (T-type); column gradients of final_norm / cat_mlp.1 and the two
// fusion-weight partials (draw[2], before the softmax backward) accumulated.
#ifndef SFB_T
#define SFB_T 128
#endif
#ifndef SFB_M
#define SFB_M 3
#endif
template <typename T, int EPL>
__global__ void __launch_bounds__(SFB_T, SFB_M) sf_post_bwd_kernel(SfP p, const float* __restrict__ dout, const float* __restrict__ cat_stats,
                                                          const float* __restrict__ fin_stats, float* __restrict__ dT,
                                                          float* __restrict__ dR, T* __restrict__ dglin, T* __restrict__ dcpre,
                                                          float* __restrict__ d_fin_g, float* __restrict__ d_fin_b,
                                                          float* __restrict__ d_cat_g, float* __restrict__ d_cat_b,
                                                          float* __restrict__ draw) {
  QV_PDL_ENTRY();
  __shared__ float red[4][8][256];
  __shared__ float sred[2][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5, c0 = lane * EPL;
  const bool act = c0 < p.C;
  const float invC = 1.f / (float)p.C;
  float w0, w1;
  softmax2(p.fw, &w0, &w1);
  float cg[EPL], cb[EPL], fg[EPL], a_fg[EPL], a_fb[EPL], a_cg[EPL], a_cb[EPL], a_w0 = 0.f, a_w1 = 0.f;
#pragma unroll
  for (int i = 0; i < EPL; ++i) cg[i] = cb[i] = fg[i] = a_fg[i] = a_fb[i] = a_cg[i] = a_cb[i] = 0.f;
  if (act) { load_vec<EPL>(p.cat_g + c0, cg); load_vec<EPL>(p.cat_b + c0, cb); load_vec<EPL>(p.fin_g + c0, fg); }
  SfRow<T, EPL> R;
  for (long row = (long)blockIdx.x * wpb + warp; row < p.rows; row += (long)gridDim.x * wpb) {
    float mean_c;
    R.compute(p, row, c0, act, cg, cb, w0, w1, cat_stats[2 * row], cat_stats[2 * row + 1], true, &mean_c);
    const float mean = fin_stats[2 * row], rstd = fin_stats[2 * row + 1];
    float dy[EPL], du[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) dy[i] = 0.f;
    if (act) load_vec<EPL>(dout + row * p.C + c0, dy);
    float s1 = 0.f, s2 = 0.f, xh[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
      xh[i] = act ? (R.u[i] - mean) * rstd : 0.f;
      a_fg[i] += dy[i] * xh[i];
      a_fb[i] += dy[i];
      du[i] = dy[i] * fg[i];
      s1 += du[i];
      s2 += du[i] * xh[i];
    }
    const float m1 = warp_sum(s1) * invC, m2 = warp_sum(s2) * invC;
    float o_t[EPL], o_r[EPL], o_g[EPL], gc[EPL];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
      du[i] = rstd * (du[i] - m1 - xh[i] * m2);
      a_w0 += du[i] * (R.t[i] + R.gate[i] * R.r[i]);
      a_w1 += du[i] * (R.t[i] + R.m[i]);
      o_t[i] = (w0 + w1) * du[i];
      o_r[i] = w0 * du[i] * R.gate[i];
      o_g[i] = w0 * du[i] * R.r[i] * R.gate[i] * (1.f - R.gate[i]);
      const float dl = w1 * du[i] * R.dsc[i] * gelu_grad_t<T>(R.lnc[i]);
      a_cg[i] += dl * R.xc[i];
      a_cb[i] += dl;
      gc[i] = dl * cg[i];
      c1 += gc[i];
      c2 += gc[i] * R.xc[i];
    }
    const float n1 = warp_sum(c1) * invC, n2 = warp_sum(c2) * invC;
#pragma unroll
    for (int i = 0; i < EPL; ++i) gc[i] = R.rstd_c * (gc[i] - n1 - R.xc[i] * n2);
    if (act) {
      store_vec<EPL>(dT + row * p.C + c0, o_t);
      store_vec<EPL>(dR + row * p.C + c0, o_r);
      store_vec<EPL>(dglin + row * p.C + c0, o_g);
      store_vec<EPL>(dcpre + row * p.C + c0, gc);
    }
  }
#pragma unroll
  for (int i = 0; i < EPL; ++i) {
    red[0][warp][lane * EPL + i] = a_fg[i]; red[1][warp][lane * EPL + i] = a_fb[i];
    red[2][warp][lane * EPL + i] = a_cg[i]; red[3][warp][lane * EPL + i] = a_cb[i];
  }
  a_w0 = warp_sum(a_w0); a_w1 = warp_sum(a_w1);
  if (lane == 0) { sred[0][warp] = a_w0; sred[1][warp] = a_w1; }
  __syncthreads();
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int w = 0; w < wpb; ++w) { a0 += red[0][w][c]; a1 += red[1][w][c]; a2 += red[2][w][c]; a3 += red[3][w][c]; }
    atomicAdd(d_fin_g + c, a0); atomicAdd(d_fin_b + c, a1); atomicAdd(d_cat_g + c, a2); atomicAdd(d_cat_b + c, a3);
  }
  if (threadIdx.x < 2) {
    float a = 0.f;
    for (int w = 0; w < wpb; ++w) a += sred[threadIdx.x][w];
    atomicAdd(draw + threadIdx.x, a * (threadIdx.x == 0 ? w0 : w1));   // w_i * dL/dw_i: what fusion_bwd_final expects
  }
}

// K1 backward: ds = LN_gate-backward(dg_in);  dT = dT_part + ds + dcat[:, :C];  dR = dR_part + ds + dcat[:, C:]
template <typename T, int EPL>
__global__ void __launch_bounds__(256, SFP_M) sf_pre_bwd_kernel(const float* __restrict__ Tin, const float* __restrict__ R,
                                                         const T* __restrict__ dg_in, const T* __restrict__ dcat, long rows, int C,
                                                         const float* __restrict__ gamma, const float* __restrict__ stats,
                                                         float* dT, float* dR, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  QV_PDL_ENTRY();
  __shared__ float red[2][8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5, c0 = lane * EPL;
  const bool act = c0 < C;
  const float invC = 1.f / (float)C;
  float gm[EPL], ag[EPL], ab[EPL];
#pragma unroll
  for (int i = 0; i < EPL; ++i) gm[i] = ag[i] = ab[i] = 0.f;
  if (act) load_vec<EPL>(gamma + c0, gm);
  for (long row = (long)blockIdx.x * wpb + warp; row < rows; row += (long)gridDim.x * wpb) {
    float t[EPL], r[EPL], dv[EPL], dc0[EPL], dc1[EPL], pt[EPL], pr[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) t[i] = r[i] = dv[i] = dc0[i] = dc1[i] = pt[i] = pr[i] = 0.f;
    if (act) {
      load_vec<EPL>(Tin + row * C + c0, t);
      load_vec<EPL>(R + row * C + c0, r);
      load_vec<EPL>(dg_in + row * C + c0, dv);
      load_vec<EPL>(dcat + row * 2 * C + c0, dc0);
      load_vec<EPL>(dcat + row * 2 * C + C + c0, dc1);
      load_vec<EPL>(dT + row * C + c0, pt);
      load_vec<EPL>(dR + row * C + c0, pr);
    }
    const float mean = stats[2 * row], rstd = stats[2 * row + 1];
    float s1 = 0.f, s2 = 0.f, xh[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
      xh[i] = act ? (t[i] + r[i] - mean) * rstd : 0.f;
      ag[i] += dv[i] * xh[i];
      ab[i] += dv[i];
      dv[i] *= gm[i];
      s1 += dv[i];
      s2 += dv[i] * xh[i];
    }
    const float m1 = warp_sum(s1) * invC, m2 = warp_sum(s2) * invC;
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
      const float ds = rstd * (dv[i] - m1 - xh[i] * m2);
      pt[i] += ds + dc0[i];
      pr[i] += ds + dc1[i];
    }
    if (act) { store_vec<EPL>(dT + row * C + c0, pt); store_vec<EPL>(dR + row * C + c0, pr); }
  }
#pragma unroll
  for (int i = 0; i < EPL; ++i) { red[0][warp][lane * EPL + i] = ag[i]; red[1][warp][lane * EPL + i] = ab[i]; }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < wpb; ++w) { a += red[0][w][c]; b += red[1][w][c]; }
    atomicAdd(dgamma + c, a);
    atomicAdd(dbeta + c, b);
  }
}

__global__ void rng_snapshot_advance_kernel(unsigned long long* rng, unsigned long long* snap) {
  QV_PDL_ENTRY();
  snap[0] = rng[0];
  snap[1] = rng[1];
  rng[1] += 1;
}

int row_grid(long rows, int mult) { return (int)max(1L, min((rows + 7) / 8, (long)qv_num_sms() * mult)); }
int vec_grid(long nvec) { return (int)max(1L, min((nvec + 255) / 256, (long)qv_num_sms() * 8)); }

}  // namespace

// ================================================================================================ launchers
#define DT_SWITCH(dt, CALL_F32, CALL_BF16) do { if ((dt) == QV_BF16) { CALL_BF16; } else { CALL_F32; } } while (0)

int bn_fwd(cudaStream_t s, int dt, const void* x, long rows, int C, const float* gamma, const float* beta, float eps,
           float momentum, int train, float* running_mean, float* running_var, long long* num_batches, int gelu,
           float* sums_scratch, float* mr, void* y) {
  if (rows <= 0) return 0;
  QV_CHECK(C % 8 == 0 && C <= 1024 && 256 % (C / 8) == 0, "batch_norm: C=%d unsupported (multiple of 8, C/8 divides 256)", C);
  if (train) {
    unsigned long long* fix = reinterpret_cast<unsigned long long*>(sums_scratch);   // 2C 64-bit sums in the 4096-float scratch
    QV_CUDA(cudaMemsetAsync(fix, 0, 2 * C * sizeof(unsigned long long), s));
    const int rpp = 256 / (C / 8);
    const int grid = (int)max(1L, min((rows + rpp - 1) / rpp, (long)qv_num_sms() * 4));
    DT_SWITCH(dt, (qv_launch(bn_stats_kernel<float>, grid, 256, 2 * C * sizeof(unsigned long long), s, (const float*)x, rows, C, fix)),
              (qv_launch(bn_stats_kernel<bf16>, grid, 256, 2 * C * sizeof(unsigned long long), s, (const bf16*)x, rows, C, fix)));
    QV_LAUNCH_CHECK();
  } else {
    QV_CHECK(running_mean && running_var, "batch_norm: eval mode needs running statistics");
  }
  DT_SWITCH(dt, (qv_launch(bn_finalize_kernel<float>, cdiv(C, 128), 128, 0, s, (const float*)x, sums_scratch, rows, C, eps, momentum, train,
                                                                        running_mean, running_var, num_batches, mr)),
            (qv_launch(bn_finalize_kernel<bf16>, cdiv(C, 128), 128, 0, s, (const bf16*)x, sums_scratch, rows, C, eps, momentum, train,
                                                                   running_mean, running_var, num_batches, mr)));
  QV_LAUNCH_CHECK();
  const long nvec = rows * C / 8;
  const int grid = vec_grid(nvec);
  if (gelu) DT_SWITCH(dt, (qv_launch(bn_apply_kernel<float, true>, grid, 256, 0, s, (const float*)x, nvec, C, mr, gamma, beta, (float*)y)),
                      (qv_launch(bn_apply_kernel<bf16, true>, grid, 256, 0, s, (const bf16*)x, nvec, C, mr, gamma, beta, (bf16*)y)));
  else DT_SWITCH(dt, (qv_launch(bn_apply_kernel<float, false>, grid, 256, 0, s, (const float*)x, nvec, C, mr, gamma, beta, (float*)y)),
                 (qv_launch(bn_apply_kernel<bf16, false>, grid, 256, 0, s, (const bf16*)x, nvec, C, mr, gamma, beta, (bf16*)y)));
  QV_LAUNCH_CHECK();
  return 0;
}

int bn_bwd(cudaStream_t s, int dt, const void* x, const void* dy, long rows, int C, const float* gamma, const float* beta,
           const float* mr, int train, int gelu, float* sums_scratch, void* dx, float* dgamma, float* dbeta) {
  if (rows <= 0) return 0;
  QV_CUDA(cudaMemsetAsync(sums_scratch, 0, 2 * C * sizeof(float), s));
  const int rpp = 256 / (C / 8);
  const int g1 = (int)max(1L, min((rows + rpp - 1) / rpp, (long)qv_num_sms() * 4));
  const size_t sm = 2 * C * sizeof(float);
#define BN_R(T, G) qv_launch(bn_bwd_reduce_kernel<T, G>, g1, 256, sm, s, (const T*)x, (const T*)dy, rows, C, mr, gamma, beta, sums_scratch)
  if (gelu) DT_SWITCH(dt, (BN_R(float, true)), (BN_R(bf16, true)));
  else DT_SWITCH(dt, (BN_R(float, false)), (BN_R(bf16, false)));
#undef BN_R
  QV_LAUNCH_CHECK();
  const long nvec = rows * C / 8;
  const int g2 = vec_grid(nvec);
#define BN_A(T, G) \
  qv_launch(bn_bwd_apply_kernel<T, G>, g2, 256, 0, s, (const T*)x, (const T*)dy, nvec, rows, C, mr, gamma, beta, sums_scratch, train, (T*)dx, dgamma, dbeta)
  if (gelu) DT_SWITCH(dt, (BN_A(float, true)), (BN_A(bf16, true)));
  else DT_SWITCH(dt, (BN_A(float, false)), (BN_A(bf16, false)));
#undef BN_A
  QV_LAUNCH_CHECK();
  return 0;
}

int im2col_img(cudaStream_t s, int dt, const float* img, int B, int Cin, int S, int Kp, void* col) {
  QV_CHECK(Kp == 32 && Cin <= 3, "im2col_img: in_channels=%d (Kp=%d) not instantiated (<= 3 channels)", Cin, Kp);
  const long rows = (long)B * (S / 2) * (S / 2);
  const int grid = vec_grid(rows);
  if (Cin == 3) DT_SWITCH(dt, (qv_launch(im2col_img_kernel<float, 4, 3>, grid, 256, 0, s, img, B, Cin, S, (float*)col)),
                          (qv_launch(im2col_img_kernel<bf16, 4, 3>, grid, 256, 0, s, img, B, Cin, S, (bf16*)col)));
  else DT_SWITCH(dt, (qv_launch(im2col_img_kernel<float, 4, 0>, grid, 256, 0, s, img, B, Cin, S, (float*)col)),
                 (qv_launch(im2col_img_kernel<bf16, 4, 0>, grid, 256, 0, s, img, B, Cin, S, (bf16*)col)));
  QV_LAUNCH_CHECK();
  return 0;
}
int im2col_nhwc(cudaStream_t s, int dt, const void* x, int B, int Hi, int Cin, void* col) {
  QV_CHECK(Cin % 8 == 0 && Hi % 2 == 0, "im2col: Cin=%d Hi=%d unsupported", Cin, Hi);
  const long total = (long)B * (Hi / 2) * (Hi / 2) * 9 * (Cin / 8);
  const int grid = vec_grid(total);
  DT_SWITCH(dt, (qv_launch(im2col_nhwc_kernel<float>, grid, 256, 0, s, (const float*)x, B, Hi, Cin, (float*)col)),
            (qv_launch(im2col_nhwc_kernel<bf16>, grid, 256, 0, s, (const bf16*)x, B, Hi, Cin, (bf16*)col)));
  QV_LAUNCH_CHECK();
  return 0;
}
int col2im_nhwc(cudaStream_t s, int dt, const void* dcol, int B, int Hi, int Cin, void* dx) {
  const long total = (long)B * Hi * Hi * (Cin / 8);
  const int grid = vec_grid(total);
  DT_SWITCH(dt, (qv_launch(col2im_nhwc_kernel<float>, grid, 256, 0, s, (const float*)dcol, B, Hi, Cin, (float*)dx)),
            (qv_launch(col2im_nhwc_kernel<bf16>, grid, 256, 0, s, (const bf16*)dcol, B, Hi, Cin, (bf16*)dx)));
  QV_LAUNCH_CHECK();
  return 0;
}
int conv_w_pack(cudaStream_t s, const float* W, int N, int Cin, int Kp, float* Wp) {
  qv_launch(conv_w_pack_kernel, cdiv((long)N * Kp, 256), 256, 0, s, W, N, Cin, Kp, Wp);
  QV_LAUNCH_CHECK();
  return 0;
}
int conv_w_unpack_add(cudaStream_t s, const float* dWp, int N, int Cin, int Kp, float* dW) {
  qv_launch(conv_w_unpack_add_kernel, cdiv((long)N * Cin * 9, 256), 256, 0, s, dWp, N, Cin, Kp, dW);
  QV_LAUNCH_CHECK();
  return 0;
}

static int row_epl(int C) { return C > 128 ? 8 : 4; }
// x: dt; y / resid: dt_y (fp32 or the activation type); y32: optional extra fp32 copy of y
int rowln_fwd(cudaStream_t s, int dt, const void* x, long rows, int C, const float* gamma, const float* beta, float eps,
              int gelu_out, int dt_y, const void* resid, const float* scale, void* y, float* y32, float* stats) {
  if (rows <= 0) return 0;
  const int epl = row_epl(C);
  QV_CHECK(C <= 256 && C % epl == 0, "rowln_fwd: C=%d unsupported", C);
  QV_CHECK(!(dt == QV_F32 && dt_y == QV_BF16), "rowln_fwd: fp32 input with bf16 output is not instantiated");
  const int grid = row_grid(rows, 8);
#define RL_F(TX, TY, E) \
  qv_launch(rowln_fwd_kernel<TX, TY, E>, grid, 256, 0, s, (const TX*)x, rows, C, gamma, beta, eps, gelu_out, (const TY*)resid, scale, (TY*)y, y32, stats)
#define RL_F2(TX, TY) do { if (epl == 8) RL_F(TX, TY, 8); else RL_F(TX, TY, 4); } while (0)
  if (dt == QV_BF16 && dt_y == QV_BF16) RL_F2(bf16, bf16);
  else if (dt == QV_BF16) RL_F2(bf16, float);
  else RL_F2(float, float);
#undef RL_F2
#undef RL_F
  QV_LAUNCH_CHECK();
  return 0;
}

// x / dx: dt; dy: dt_dy
int rowln_bwd(cudaStream_t s, int dt, const void* x, int dt_dy, const void* dy, long rows, int C, const float* gamma, const float* beta,
              const float* stats, int gelu_out, const float* scale, float* dscale, void* dx, float* dgamma, float* dbeta) {
  if (rows <= 0) return 0;
  const int epl = row_epl(C);
  QV_CHECK(C <= 256 && C % epl == 0, "rowln_bwd: C=%d unsupported", C);
  QV_CHECK(!(dt == QV_F32 && dt_dy == QV_BF16), "rowln_bwd: fp32 activations with bf16 gradients are not instantiated");
  const int grid = row_grid((rows + 1) / 2, 6);
#define RL_B(TX, TD, E) \
  qv_launch(rowln_bwd_kernel<TX, TD, E>, grid, 256, 0, s, (const TX*)x, (const TD*)dy, rows, C, gamma, beta, stats, gelu_out, scale, dscale, (TX*)dx, dgamma, dbeta)
#define RL_B2(TX, TD) do { if (epl == 8) RL_B(TX, TD, 8); else RL_B(TX, TD, 4); } while (0)
  if (dt == QV_BF16 && dt_dy == QV_BF16) RL_B2(bf16, bf16);
  else if (dt == QV_BF16) RL_B2(bf16, float);
  else RL_B2(float, float);
#undef RL_B2
#undef RL_B
  QV_LAUNCH_CHECK();
  return 0;
}

int sf_pre_fwd(cudaStream_t s, int dt, const float* Tin, const float* R, long rows, int C, const float* gamma, const float* beta,
               void* g_in, void* cat, float* stats) {
  const int epl = row_epl(C);
  QV_CHECK(C <= 256 && C % epl == 0, "splitfusion: C=%d unsupported", C);
  const int grid = row_grid(rows, 8);
#define SF_GO(T, E) qv_launch(sf_pre_fwd_kernel<T, E>, grid, 256, 0, s, Tin, R, rows, C, gamma, beta, (T*)g_in, (T*)cat, stats)
  if (dt == QV_BF16) { if (epl == 8) SF_GO(bf16, 8); else SF_GO(bf16, 4); }
  else { if (epl == 8) SF_GO(float, 8); else SF_GO(float, 4); }
#undef SF_GO
  QV_LAUNCH_CHECK();
  return 0;
}

int sf_post_fwd(cudaStream_t s, int dt, const SfArgs& a, float* out, float* cat_stats, float* fin_stats) {
  const int epl = row_epl(a.C);
  SfP p{a.Tin, a.R, a.glin, a.cpre, a.cat_g, a.cat_b, a.fin_g, a.fin_b, a.fw, a.drop_p, a.rng, a.site, a.rows, a.C};
  const int grid = row_grid(a.rows, 8);
#define SF_GO(T, E) qv_launch(sf_post_fwd_kernel<T, E>, grid, 256, 0, s, p, out, cat_stats, fin_stats)
  if (dt == QV_BF16) { if (epl == 8) SF_GO(bf16, 8); else SF_GO(bf16, 4); }
  else { if (epl == 8) SF_GO(float, 8); else SF_GO(float, 4); }
#undef SF_GO
  QV_LAUNCH_CHECK();
  return 0;
}

int sf_post_bwd(cudaStream_t s, int dt, const SfArgs& a, const float* dout, const float* cat_stats, const float* fin_stats, float* dT,
                float* dR, void* dglin, void* dcpre, float* d_fin_g, float* d_fin_b, float* d_cat_g, float* d_cat_b, float* draw) {
  const int epl = row_epl(a.C);
  SfP p{a.Tin, a.R, a.glin, a.cpre, a.cat_g, a.cat_b, a.fin_g, a.fin_b, a.fw, a.drop_p, a.rng, a.site, a.rows, a.C};
  const int grid = row_grid((a.rows + 1) / 2, 4 * (256 / SFB_T));
#define SF_GO(T, E) \
  qv_launch(sf_post_bwd_kernel<T, E>, grid, SFB_T, 0, s, p, dout, cat_stats, fin_stats, dT, dR, (T*)dglin, (T*)dcpre, d_fin_g, d_fin_b, d_cat_g, d_cat_b, draw)
  if (dt == QV_BF16) { if (epl == 8) SF_GO(bf16, 8); else SF_GO(bf16, 4); }
  else { if (epl == 8) SF_GO(float, 8); else SF_GO(float, 4); }
#undef SF_GO
  QV_LAUNCH_CHECK();
  return 0;
}

int sf_pre_bwd(cudaStream_t s, int dt, const float* Tin, const float* R, const void* dg_in, const void* dcat, long rows, int C,
               const float* gamma, const float* stats, float* dT, float* dR, float* dgamma, float* dbeta) {
  const int epl = row_epl(C);
  const int grid = row_grid((rows + 1) / 2, 6);
#define SF_GO(T, E) \
  qv_launch(sf_pre_bwd_kernel<T, E>, grid, 256, 0, s, Tin, R, (const T*)dg_in, (const T*)dcat, rows, C, gamma, stats, dT, dR, dgamma, dbeta)
  if (dt == QV_BF16) { if (epl == 8) SF_GO(bf16, 8); else SF_GO(bf16, 4); }
  else { if (epl == 8) SF_GO(float, 8); else SF_GO(float, 4); }
#undef SF_GO
  QV_LAUNCH_CHECK();
  return 0;
}

int rng_snapshot_advance(cudaStream_t s, unsigned long long* rng, unsigned long long* snap) {
  qv_launch(rng_snapshot_advance_kernel, 1, 1, 0, s, rng, snap);
  QV_LAUNCH_CHECK();
  return 0;
}

// =============================================================================== bilinear resize of a channels-last map
// F.interpolate(x, size = (Ho, Wo), mode = 'bilinear', align_corners = False) of LMFAdapter (H:839-843): taken when a model built
// for 32 x 32 images is fed larger ones (STL-10 recipe, 96 x 96: 24 x 24 maps -> 8 x 8; the source index is 3 i + 1 exactly, so the
// resize degenerates to a strided pick there -- zero-weight taps are skipped).  Source index as in ATen's
// area_pixel_compute_source_index: src = max(0, scale (dst + 0.5) - 0.5), scale = in / out.
namespace {
struct Tap { int i0, i1; float w0, w1; };
__device__ __forceinline__ Tap make_tap(int o, int in, int out) {
  const float scale = (float)in / (float)out;
  float src = scale * ((float)o + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  Tap t;
  t.i0 = min((int)src, in - 1);
  t.i1 = min(t.i0 + 1, in - 1);
  t.w1 = src - (float)t.i0;
  t.w0 = 1.f - t.w1;
  return t;
}
template <typename T>
__global__ void __launch_bounds__(256) resize_bilinear_fwd_kernel(const T* __restrict__ in, int B, int Hi, int Wi, int Ho, int Wo, int C,
                                                                  T* __restrict__ out) {
  QV_PDL_ENTRY();
  const int c2n = C / 2;
  const long total = (long)B * Ho * Wo * c2n;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % c2n) * 2;
    long r = idx / c2n;
    const int ox = (int)(r % Wo); r /= Wo;
    const int oy = (int)(r % Ho);
    const long b = r / Ho;
    const Tap ty = make_tap(oy, Hi, Ho), tx = make_tap(ox, Wi, Wo);
    const T* base = in + b * (long)Hi * Wi * C + c;
    const float2 a = ld2(base + ((long)ty.i0 * Wi + tx.i0) * C), bb = ld2(base + ((long)ty.i0 * Wi + tx.i1) * C);
    const float2 cc = ld2(base + ((long)ty.i1 * Wi + tx.i0) * C), dd = ld2(base + ((long)ty.i1 * Wi + tx.i1) * C);
    float2 v;
    v.x = ty.w0 * (tx.w0 * a.x + tx.w1 * bb.x) + ty.w1 * (tx.w0 * cc.x + tx.w1 * dd.x);
    v.y = ty.w0 * (tx.w0 * a.y + tx.w1 * bb.y) + ty.w1 * (tx.w0 * cc.y + tx.w1 * dd.y);
    st2(out + idx * 2, v);
  }
}
// din (pre-zeroed) += taps * dout.  A thread owns two channels of one image and walks the output pixels sequentially: several
// output pixels can hit the same source pixel, this order has no write races and is deterministic.
template <typename T>
__global__ void __launch_bounds__(256) resize_bilinear_bwd_kernel(const T* __restrict__ dout, int B, int Hi, int Wi, int Ho, int Wo, int C,
                                                                  T* __restrict__ din) {
  QV_PDL_ENTRY();
  const int c2n = C / 2;
  const long total = (long)B * c2n;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % c2n) * 2;
    const long b = idx / c2n;
    T* base = din + b * (long)Hi * Wi * C + c;
    for (int oy = 0; oy < Ho; ++oy) {
      const Tap ty = make_tap(oy, Hi, Ho);
      for (int ox = 0; ox < Wo; ++ox) {
        const Tap tx = make_tap(ox, Wi, Wo);
        const float2 g = ld2(dout + ((b * Ho + oy) * (long)Wo + ox) * C + c);
        const int ys[2] = {ty.i0, ty.i1}, xs[2] = {tx.i0, tx.i1};
        const float wy[2] = {ty.w0, ty.w1}, wx[2] = {tx.w0, tx.w1};
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const float w = wy[i] * wx[j];
            if (w == 0.f) continue;
            T* p = base + ((long)ys[i] * Wi + xs[j]) * C;
            float2 o = ld2(p);
            o.x += w * g.x; o.y += w * g.y;
            st2(p, o);
          }
      }
    }
  }
}
}  // namespace

int resize_bilinear_fwd(cudaStream_t s, int dt, const void* in, int B, int Hi, int Wi, int Ho, int Wo, int C, void* out) {
  QV_CHECK(C % 2 == 0, "resize: channel count must be even");
  const long total = (long)B * Ho * Wo * (C / 2);
  if (total <= 0) return 0;
  const int grid = (int)min((long)qv_num_sms() * 16, (total + 255) / 256);
  DT_SWITCH(dt, (qv_launch(resize_bilinear_fwd_kernel<float>, grid, 256, 0, s, (const float*)in, B, Hi, Wi, Ho, Wo, C, (float*)out)),
            (qv_launch(resize_bilinear_fwd_kernel<bf16>, grid, 256, 0, s, (const bf16*)in, B, Hi, Wi, Ho, Wo, C, (bf16*)out)));
  QV_LAUNCH_CHECK();
  return 0;
}
int resize_bilinear_bwd(cudaStream_t s, int dt, const void* dout, int B, int Hi, int Wi, int Ho, int Wo, int C, void* din) {
  QV_CHECK(C % 2 == 0, "resize: channel count must be even");
  const long total = (long)B * (C / 2);
  if (total <= 0) return 0;
  QV_CUDA(cudaMemsetAsync(din, 0, (size_t)B * Hi * Wi * C * (dt == QV_BF16 ? 2 : 4), s));
  const int grid = (int)min((long)qv_num_sms() * 16, (total + 255) / 256);
  DT_SWITCH(dt, (qv_launch(resize_bilinear_bwd_kernel<float>, grid, 256, 0, s, (const float*)dout, B, Hi, Wi, Ho, Wo, C, (float*)din)),
            (qv_launch(resize_bilinear_bwd_kernel<bf16>, grid, 256, 0, s, (const bf16*)dout, B, Hi, Wi, Ho, Wo, C, (bf16*)din)));
  QV_LAUNCH_CHECK();
  return 0;
}
