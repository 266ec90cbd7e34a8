// Gradient clipping (per-parameter + global L2, H:1413-1432) and the fused AdamW step (H:1436, torch single-tensor
// rule) on flat fp32 buffers.  HBM-bound: AdamW reads p, g, m, v (16 B/elem) and writes p, m, v (12 B/elem).
// Segments (= parameter tensors) start on 4-float boundaries so every access is a float4.
#include "../../include/qavit_b200.h"
#include "kernels.h"

namespace {

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x < 32) {
    r = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    r = warp_sum(r);
  }
  return r;  // valid in thread 0
}

// norms[s] = ||g_s||_2, one CTA per segment
__global__ void __launch_bounds__(256) seg_norm_kernel(const float* __restrict__ g, const long long* __restrict__ off,
                                                       const int* __restrict__ flags, float* __restrict__ norms) {
  QV_PDL_ENTRY();
  __shared__ float red[32];
  const int s = blockIdx.x;
  float acc = 0.f;
  if (flags[s] & 1) {
    const long long b = off[s], e = off[s + 1];
    for (long long i = b + threadIdx.x * 4LL; i < e; i += blockDim.x * 4LL) {
      const float4 v = *reinterpret_cast<const float4*>(g + i);
      acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
  }
  const float t = block_sum(acc, red);
  if (threadIdx.x == 0) norms[s] = sqrtf(t);
}

// per-segment scale = per-param coef * global coef (written over norms[s]); norms[n] = global norm before global clip
// prescale: a factor every gradient carries before clipping (1 / world_size when the buffer holds the SUM over data-parallel
// ranks): norms are taken of prescale * g and the factor is folded into the per-segment scale, so no separate division pass.
__global__ void __launch_bounds__(1024) clip_coef_kernel(const int* __restrict__ flags, int n, float per_param_max,
                                                         float max_norm, float prescale, float* __restrict__ norms) {
  QV_PDL_ENTRY();
  __shared__ float red[32];
  __shared__ float gcoef;
  float acc = 0.f;
  for (int s = threadIdx.x; s < n; s += blockDim.x) {
    float c = 1.f;
    const float nm = norms[s] * prescale;
    if (flags[s] & 2) c = fminf(1.f, per_param_max / (nm + 1e-6f));
    const float cn = (flags[s] & 1) ? c * nm : 0.f;
    acc += cn * cn;
    norms[s] = c;
  }
  const float t = block_sum(acc, red);
  if (threadIdx.x == 0) {
    const float total = sqrtf(t);
    norms[n] = total;
    gcoef = fminf(1.f, max_norm / (total + 1e-6f));
    norms[n + 1] = gcoef;
  }
  __syncthreads();
  for (int s = threadIdx.x; s < n; s += blockDim.x) norms[s] *= gcoef * prescale;
}
__global__ void scaled_copy_kernel(const float* __restrict__ x, float scale, long n, float* __restrict__ y) {
  QV_PDL_ENTRY();
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) y[i] = x[i] * scale;
}

__device__ __forceinline__ int find_seg(const long long* __restrict__ off, int n, long long i) {
  int lo = 0, hi = n;  // off[lo] <= i < off[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (off[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256) scale_grads_kernel(float* __restrict__ g, const long long* __restrict__ off,
                                                          const int* __restrict__ flags, int n,
                                                          const float* __restrict__ scale, long long total) {
  QV_PDL_ENTRY();
  for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4; i < total; i += (long long)gridDim.x * blockDim.x * 4) {
    const int s = find_seg(off, n, i);
    if (!(flags[s] & 1)) continue;
    const float c = scale[s];
    float4 v = *reinterpret_cast<float4*>(g + i);
    v.x *= c; v.y *= c; v.z *= c; v.w *= c;
    *reinterpret_cast<float4*>(g + i) = v;
  }
}

// hyper: lr, beta1, beta2, eps, wd, 1 - beta1^t, 1 - beta2^t, ema decay d, 1 - d (formed in double on the host, like torch's alpha)
// EMA: ModelEMA.update (H:139-149) folded in -- ema = d * ema + (1 - d) * p_new for EVERY parameter (also the ones AdamW
// skips), one extra fp32 stream through the same pass instead of 815 mul_/add_ launches.
template <bool EMA>
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, float* __restrict__ ema, const long long* __restrict__ off,
                                                    const int* __restrict__ flags, int n, const float* __restrict__ hyper,
                                                    long long total) {
  QV_PDL_ENTRY();
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4], bc1 = hyper[5], bc2 = hyper[6];
  const float step = lr / bc1, rs2 = rsqrtf(bc2), decay = 1.f - lr * wd;
  const float ed = EMA ? hyper[7] : 0.f, ec = EMA ? hyper[8] : 0.f;
  for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4; i < total; i += (long long)gridDim.x * blockDim.x * 4) {
    const int s = find_seg(off, n, i);
    if (!(flags[s] & 1)) {           // grad is None: skipped entirely, also by weight decay
      if (EMA) {
        const float4 pp = *reinterpret_cast<const float4*>(p + i);
        float4 ee = *reinterpret_cast<float4*>(ema + i);
        ee.x = __fadd_rn(__fmul_rn(ee.x, ed), __fmul_rn(pp.x, ec)); ee.y = __fadd_rn(__fmul_rn(ee.y, ed), __fmul_rn(pp.y, ec));
        ee.z = __fadd_rn(__fmul_rn(ee.z, ed), __fmul_rn(pp.z, ec)); ee.w = __fadd_rn(__fmul_rn(ee.w, ed), __fmul_rn(pp.w, ec));
        *reinterpret_cast<float4*>(ema + i) = ee;
      }
      continue;
    }
    float4 pp = *reinterpret_cast<float4*>(p + i);
    const float4 gg = *reinterpret_cast<const float4*>(g + i);
    float4 mm = *reinterpret_cast<float4*>(m + i);
    float4 vv = *reinterpret_cast<float4*>(v + i);
#define ADAM1(P, G, M, V)                              \
  P *= decay;                                          \
  M = b1 * M + (1.f - b1) * G;                         \
  V = b2 * V + (1.f - b2) * G * G;                     \
  P -= step * M / (sqrtf(V) * rs2 + eps);
    ADAM1(pp.x, gg.x, mm.x, vv.x)
    ADAM1(pp.y, gg.y, mm.y, vv.y)
    ADAM1(pp.z, gg.z, mm.z, vv.z)
    ADAM1(pp.w, gg.w, mm.w, vv.w)
#undef ADAM1
    *reinterpret_cast<float4*>(p + i) = pp;
    *reinterpret_cast<float4*>(m + i) = mm;
    *reinterpret_cast<float4*>(v + i) = vv;
    if (EMA) {   // mul_(decay).add_(p, alpha = 1 - decay): two roundings per term, no fma contraction (bit-exact vs torch)
      float4 ee = *reinterpret_cast<float4*>(ema + i);
      ee.x = __fadd_rn(__fmul_rn(ee.x, ed), __fmul_rn(pp.x, ec)); ee.y = __fadd_rn(__fmul_rn(ee.y, ed), __fmul_rn(pp.y, ec));
      ee.z = __fadd_rn(__fmul_rn(ee.z, ed), __fmul_rn(pp.z, ec)); ee.w = __fadd_rn(__fmul_rn(ee.w, ed), __fmul_rn(pp.w, ec));
      *reinterpret_cast<float4*>(ema + i) = ee;
    }
  }
}

// CutMix / MixUp of a batch against a permutation of itself (H:1379-1399).  mode 1: pixels inside [x1, x2) x [y1, y2) come
// from image perm[b]; mode 2: lam * x[b] + (1 - lam) * x[perm[b]] (two roundings like the torch expression).
__global__ void __launch_bounds__(256) batch_mix_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                        const long long* __restrict__ perm, int B, int C, int H, int W, int mode,
                                                        int x1, int y1, int x2, int y2, float lam, float lb) {
  QV_PDL_ENTRY();
  const long long per = (long long)C * H * W, total = (long long)B * per;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / per, r = i % per;
    const int x = (int)(r % W), y = (int)((r / W) % H);
    const long long j = perm[b] * per + r;
    float v = in[i];
    if (mode == 1) {
      if (x >= x1 && x < x2 && y >= y1 && y < y2) v = in[j];
    } else if (mode == 2) {
      v = __fadd_rn(__fmul_rn(lam, v), __fmul_rn(lb, in[j]));
    }
    out[i] = v;
  }
}

// transforms.ToTensor() + transforms.Normalize(mean, std) (H:1300, HQAViT_C100_Finetune.py:100-103) on a whole batch:
// out[b, c, y, x] = (in / 255 - mean[c]) / std[c], one rounding per torch op (div, sub, div) so the result is bit-identical to the
// torchvision pipeline; hflip = RandomHorizontalFlip(p = 1) of the TTA views.  in_kind 0: uint8 [B, H, W, C] (the datasets' raw
// layout), 1: uint8 [B, C, H, W], 2: float [B, C, H, W] already in [0, 1] (only Normalize is applied).
__global__ void __launch_bounds__(256) normalize_images_kernel(const void* __restrict__ in, int in_kind, int B, int C, int H, int W,
                                                               const float* __restrict__ mean, const float* __restrict__ stdv, int hflip,
                                                               float* __restrict__ out) {
  QV_PDL_ENTRY();
  const long long total = (long long)B * C * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H), c = (int)((i / ((long long)W * H)) % C);
    const long long b = i / ((long long)W * H * C);
    const int xs = hflip ? W - 1 - x : x;
    float v;
    if (in_kind == 0) v = __fdiv_rn((float)static_cast<const unsigned char*>(in)[((b * H + y) * W + xs) * C + c], 255.f);
    else if (in_kind == 1) v = __fdiv_rn((float)static_cast<const unsigned char*>(in)[((b * C + c) * H + y) * W + xs], 255.f);
    else v = static_cast<const float*>(in)[((b * C + c) * H + y) * W + xs];
    out[i] = __fdiv_rn(__fsub_rn(v, mean[c]), stdv[c]);
  }
}

}  // namespace

extern "C" int qavit_normalize_images(const void* in, int in_kind, int B, int C, int H, int W, const float* mean, const float* stdv,
                                      int hflip, float* out, void* stream) {
  QV_CHECK(in && mean && stdv && out && (const void*)in != (const void*)out, "normalize_images: null / aliased argument");
  QV_CHECK(in_kind >= 0 && in_kind <= 2, "normalize_images: in_kind %d (0 = uint8 NHWC, 1 = uint8 NCHW, 2 = float NCHW)", in_kind);
  const long long total = (long long)B * C * H * W;
  if (total <= 0) return 0;
  const int grid = (int)max(1LL, min((long long)qv_num_sms() * 16, (total + 255) / 256));
  qv_launch(normalize_images_kernel, grid, 256, 0, (cudaStream_t)stream, in, in_kind, B, C, H, W, mean, stdv, hflip, out);
  QV_LAUNCH_CHECK();
  return 0;
}

extern "C" int qavit_scaled_copy(const float* x, float scale, long long n, float* y, void* stream) {
  QV_CHECK(x && y, "scaled_copy: null argument");
  if (n <= 0) return 0;
  qv_launch(scaled_copy_kernel, (int)max(1LL, min((long long)qv_num_sms() * 8, (n + 255) / 256)), 256, 0, (cudaStream_t)stream, x, scale, (long)n, y);
  QV_LAUNCH_CHECK();
  return 0;
}

extern "C" int qavit_clip_grads(float* grads, const long long* seg_off, const int* seg_flags, int n_seg, float per_param_max,
                                float max_norm, float* norms, long long total, void* stream) {
  return qavit_clip_grads_scaled(grads, seg_off, seg_flags, n_seg, per_param_max, max_norm, 1.0f, norms, total, stream);
}

extern "C" int qavit_clip_grads_scaled(float* grads, const long long* seg_off, const int* seg_flags, int n_seg, float per_param_max,
                                       float max_norm, float prescale, float* norms, long long total, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (n_seg <= 0) return 0;
  qv_launch(seg_norm_kernel, n_seg, 256, 0, s, grads, seg_off, seg_flags, norms);
  QV_LAUNCH_CHECK();
  qv_launch(clip_coef_kernel, 1, 1024, 0, s, seg_flags, n_seg, per_param_max, max_norm, prescale, norms);
  QV_LAUNCH_CHECK();
  const int grid = (int)max(1LL, min((long long)qv_num_sms() * 8, (total / 4 + 255) / 256));
  qv_launch(scale_grads_kernel, grid, 256, 0, s, grads, seg_off, seg_flags, n_seg, norms, total);
  QV_LAUNCH_CHECK();
  return 0;
}

extern "C" int qavit_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const long long* seg_off,
                                const int* seg_flags, int n_seg, const float* hyper, long long total, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (n_seg <= 0) return 0;
  const int grid = (int)max(1LL, min((long long)qv_num_sms() * 8, (total / 4 + 255) / 256));
  qv_launch(adamw_kernel<false>, grid, 256, 0, s, params, grads, exp_avg, exp_avg_sq, nullptr, seg_off, seg_flags, n_seg, hyper, total);
  QV_LAUNCH_CHECK();
  return 0;
}

extern "C" int qavit_adamw_ema_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* ema,
                                    const long long* seg_off, const int* seg_flags, int n_seg, const float* hyper, long long total,
                                    void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (n_seg <= 0) return 0;
  QV_CHECK(ema, "adamw_ema_step: ema buffer missing");
  const int grid = (int)max(1LL, min((long long)qv_num_sms() * 8, (total / 4 + 255) / 256));
  qv_launch(adamw_kernel<true>, grid, 256, 0, s, params, grads, exp_avg, exp_avg_sq, ema, seg_off, seg_flags, n_seg, hyper, total);
  QV_LAUNCH_CHECK();
  return 0;
}

extern "C" int qavit_segment_norms(const float* buf, const long long* seg_off, const int* seg_flags, int n_seg, float* norms,
                                   void* stream) {
  if (n_seg <= 0) return 0;
  qv_launch(seg_norm_kernel, n_seg, 256, 0, (cudaStream_t)stream, buf, seg_off, seg_flags, norms);
  QV_LAUNCH_CHECK();
  return 0;
}

extern "C" int qavit_batch_mix(const float* in, float* out, const long long* perm, int B, int C, int H, int W, int mode, int x1,
                               int y1, int x2, int y2, float lam, float lam_b, void* stream) {
  QV_CHECK(in && out && perm && in != out, "batch_mix: null / aliased argument (out must not alias in: rows are read through perm)");
  QV_CHECK(mode == 1 || mode == 2, "batch_mix: mode %d (1 = cutmix, 2 = mixup)", mode);
  const long long total = (long long)B * C * H * W;
  if (total <= 0) return 0;
  const int grid = (int)max(1LL, min((long long)qv_num_sms() * 16, (total + 255) / 256));
  qv_launch(batch_mix_kernel, grid, 256, 0, (cudaStream_t)stream, in, out, perm, B, C, H, W, mode, x1, y1, x2, y2, lam, lam_b);
  QV_LAUNCH_CHECK();
  return 0;
}
