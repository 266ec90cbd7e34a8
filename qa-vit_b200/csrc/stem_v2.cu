// Kernels of HQAViTv2's CNN stem (scope row f-1, second half; V = HQAViTv2_CIFAR100.py):
//   - the 4x4 stride-4 patchify convolution as a patch gather + GEMM                         (V:764)
//   - "spatial" LayerNorm over a whole [C, H, W] map per image with an element-wise affine    (V:765, 776, 790)
//   - LayerScale (per-channel gamma after pwconv2, V:730-746) folded into the weights: forward and dX GEMMs run on
//     diag(gamma) W2 / gamma * b2 copies, the gradient of gamma comes out of the weight-gradient GEMM's result
//   - per-image row scaling (DropPath of the stem's ConvNeXt blocks, V:748)
// Maps are channels-last rows [B * H * W, C] like the rest of the lateral path; the LayerNorm parameters keep the
// reference's [C, H, W] layout, so element (pixel, c) of an image uses parameter c * HW + pixel.
#include "kernels.h"

namespace {

template <typename T>
__global__ void __launch_bounds__(256) patch_rows_kernel(const float* __restrict__ img, int B, int Cin, int S, int p, T* __restrict__ col) {
  QV_PDL_ENTRY();
  const int n_side = S / p, N = n_side * n_side, K = Cin * p * p;
  const long total = (long)B * N * K / 2;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
    const int k = (int)((2 * i) % K);
    const long row = (2 * i) / K;
    const int n = (int)(row % N);
    const long b = row / N;
    const int c = k / (p * p), r = (k / p) % p, q = k % p;      // p is even: (k, k + 1) share an image row
    const float2 v = *reinterpret_cast<const float2*>(img + ((b * Cin + c) * S + (n / n_side) * p + r) * S + (n % n_side) * p + q);
    st2(col + 2 * i, v);
  }
}

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5;
  __syncthreads();                       // red may still be read by the previous reduction
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += red[i];
  return t;
}

// One CTA walks images blockIdx.x, blockIdx.x + gridDim.x, ...; a thread owns VPT vectors of 8 consecutive elements of the
// E = HW * C element map.  Parameters are staged once per CTA in shared memory, permuted into the map's memory order.
template <typename T, int VPT>
__global__ void __launch_bounds__(256) sln_fwd_kernel(const T* __restrict__ x, int B, int HW, int C, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, float eps, T* __restrict__ y,
                                                      float* __restrict__ stats) {
  QV_PDL_ENTRY();
  extern __shared__ float sm[];
  const int E = HW * C;
  float* g_s = sm;
  float* b_s = sm + E;
  __shared__ float red[8];
  for (int m = threadIdx.x; m < E; m += 256) {
    const int pix = m / C, c = m % C;
    g_s[m] = gamma[c * HW + pix];
    b_s[m] = beta[c * HW + pix];
  }
  __syncthreads();
  const float invE = 1.f / (float)E;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const T* xb = x + (long)b * E;
    float v[VPT][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      load_vec<8>(xb + (i * 256 + threadIdx.x) * 8, v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];
    }
    const float mean = block_sum(s, red) * invE;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(block_sum(q, red) * invE + eps);
    if (threadIdx.x == 0) { stats[2 * b] = mean; stats[2 * b + 1] = rstd; }
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int m0 = (i * 256 + threadIdx.x) * 8;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf((v[i][j] - mean) * rstd, g_s[m0 + j], b_s[m0 + j]);
      store_vec<8>(y + (long)b * E + m0, o);
    }
  }
}

// dx = rstd * (g - mean(g) - xh * mean(g * xh)), g = dy * gamma;  dgamma += dy * xh, dbeta += dy, accumulated per CTA in shared
// memory over its images (every element has one owner thread: no atomics until the final flush)
template <typename T, int VPT>
__global__ void __launch_bounds__(256) sln_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, int B, int HW, int C,
                                                      const float* __restrict__ gamma, const float* __restrict__ stats,
                                                      const T* __restrict__ resid, T* __restrict__ dx, float* __restrict__ dgamma,
                                                      float* __restrict__ dbeta) {
  QV_PDL_ENTRY();
  extern __shared__ float sm[];
  const int E = HW * C;
  float* g_s = sm;
  float* dg_s = sm + E;
  float* db_s = sm + 2 * E;
  __shared__ float red[8];
  for (int m = threadIdx.x; m < E; m += 256) {
    g_s[m] = gamma[(m % C) * HW + m / C];
    dg_s[m] = 0.f;
    db_s[m] = 0.f;
  }
  __syncthreads();
  const float invE = 1.f / (float)E;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const float mean = stats[2 * b], rstd = stats[2 * b + 1];
    float xh[VPT][8], g[VPT][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int m0 = (i * 256 + threadIdx.x) * 8;
      float d[8];
      load_vec<8>(x + (long)b * E + m0, xh[i]);
      load_vec<8>(dy + (long)b * E + m0, d);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xh[i][j] = (xh[i][j] - mean) * rstd;
        g[i][j] = d[j] * g_s[m0 + j];
        s1 += g[i][j];
        s2 = fmaf(g[i][j], xh[i][j], s2);
        dg_s[m0 + j] = fmaf(d[j], xh[i][j], dg_s[m0 + j]);
        db_s[m0 + j] += d[j];
      }
    }
    s1 = block_sum(s1, red) * invE;
    s2 = block_sum(s2, red) * invE;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = rstd * (g[i][j] - s1 - xh[i][j] * s2);
      if (resid) {      // gradient arriving at the same map from another consumer (the stage's LMFAdapter)
        float r[8];
        load_vec<8>(resid + (long)b * E + (i * 256 + threadIdx.x) * 8, r);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += r[j];
      }
      store_vec<8>(dx + (long)b * E + (i * 256 + threadIdx.x) * 8, o);
    }
  }
  __syncthreads();
  for (int m = threadIdx.x; m < E; m += 256) {
    const int idx = (m % C) * HW + m / C;
    atomicAdd(dgamma + idx, dg_s[m]);
    atomicAdd(dbeta + idx, db_s[m]);
  }
}

// W2s = diag(gamma) W2, b2s = gamma * b2  (fp32; the bf16 copies are made by the ordinary weight conversion)
__global__ void layerscale_prepare_kernel(const float* __restrict__ W, const float* __restrict__ b, const float* __restrict__ gamma,
                                          int N, int K, float* __restrict__ Ws, float* __restrict__ bs) {
  QV_PDL_ENTRY();
  const long total = (long)N * K;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) Ws[i] = W[i] * gamma[i / K];
  if (blockIdx.x == 0)
    for (int n = threadIdx.x; n < N; n += 256) bs[n] = b[n] * gamma[n];
}
// G = d/d(W2s), gb = d/d(b2s):  dW2 += gamma_n G[n, :], db2 += gamma_n gb_n, dgamma_n += <G[n, :], W2[n, :]> + gb_n b2_n
__global__ void __launch_bounds__(256) layerscale_finish_kernel(const float* __restrict__ G, const float* __restrict__ gb,
                                                                const float* __restrict__ W, const float* __restrict__ b,
                                                                const float* __restrict__ gamma, int K, float* __restrict__ dW,
                                                                float* __restrict__ db, float* __restrict__ dgamma) {
  QV_PDL_ENTRY();
  __shared__ float red[8];
  const int n = blockIdx.x;
  const float gm = gamma[n];
  float acc = 0.f;
  for (int k = threadIdx.x; k < K; k += 256) {
    const float g = G[(long)n * K + k];
    acc = fmaf(g, W[(long)n * K + k], acc);
    dW[(long)n * K + k] += gm * g;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) {
    dgamma[n] += acc + gb[n] * b[n];
    db[n] += gm * gb[n];
  }
}

// y[r, :] = (resid ? resid[r, :] : 0) + rowscale[r / rows_per_img] * x[r, :]     (y may alias x)
template <typename T>
__global__ void __launch_bounds__(256) scale_rows_kernel(const T* __restrict__ x, long nvec, int C, const float* __restrict__ rowscale,
                                                         int rows_per_img, const T* __restrict__ resid, T* __restrict__ y) {
  QV_PDL_ENTRY();
  const int vpr = C / 8;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < nvec; i += (long)gridDim.x * 256) {
    const float rs = rowscale[(i / vpr) / rows_per_img];
    float v[8], r[8];
    load_vec<8>(x + i * 8, v);
    if (resid) {
      load_vec<8>(resid + i * 8, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaf(rs, v[j], r[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= rs;
    }
    store_vec<8>(y + i * 8, v);
  }
}

// keep scales of n_sites DropPath sites for B images each: rs[site * B + b]; rate 0 -> 1
struct DpRates { float p[7]; };
__global__ void stem_droppath_kernel(const unsigned long long* rng, uint32_t site0, DpRates rates, int n_sites, int B, float* rs) {
  QV_PDL_ENTRY();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_sites * B; i += gridDim.x * blockDim.x) {
    const int site = i / B;
    DropP d;
    d.p = rates.p[site]; d.rng = rng; d.site = site0 + (uint32_t)site;
    float v = 1.f;
    if (d.p > 0.f) { const DropState st = drop_state(d); v = drop_keep1(st, (unsigned long long)(i % B)); }
    rs[i] = v;
  }
}

int vgrid(long n) { return (int)max(1L, min((n + 255) / 256, (long)qv_num_sms() * 8)); }

}  // namespace

int patch_rows(cudaStream_t s, int dt, const float* img, int B, int Cin, int S, int p, void* col) {
  QV_CHECK(p % 2 == 0 && S % p == 0, "patch_rows: patch %d / image %d", p, S);
  const long total = (long)B * (S / p) * (S / p) * Cin * p * p / 2;
  if (dt == QV_F32) qv_launch(patch_rows_kernel<float>, vgrid(total), 256, 0, s, img, B, Cin, S, p, (float*)col);
  else qv_launch(patch_rows_kernel<bf16>, vgrid(total), 256, 0, s, img, B, Cin, S, p, (bf16*)col);
  QV_LAUNCH_CHECK();
  return 0;
}

bool sln_ok(int HW, int C) {
  const int E = HW * C;
  return C % 8 == 0 && E % 2048 == 0 && (E / 2048 == 1 || E / 2048 == 2 || E / 2048 == 4 || E / 2048 == 8);
}

#define SLN_DISPATCH(KERN, T, ...)                                                                 \
  do {                                                                                             \
    switch (E / 2048) {                                                                            \
      case 1: QV_CUDA(cudaFuncSetAttribute(KERN<T, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
              qv_launch(KERN<T, 1>, grid, 256, smem, s, __VA_ARGS__); break;                       \
      case 2: QV_CUDA(cudaFuncSetAttribute(KERN<T, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
              qv_launch(KERN<T, 2>, grid, 256, smem, s, __VA_ARGS__); break;                       \
      case 4: QV_CUDA(cudaFuncSetAttribute(KERN<T, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
              qv_launch(KERN<T, 4>, grid, 256, smem, s, __VA_ARGS__); break;                       \
      default: QV_CUDA(cudaFuncSetAttribute(KERN<T, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
              qv_launch(KERN<T, 8>, grid, 256, smem, s, __VA_ARGS__); break;                       \
    }                                                                                              \
  } while (0)

int sln_fwd(cudaStream_t s, int dt, const void* x, int B, int HW, int C, const float* gamma, const float* beta, float eps, void* y,
            float* stats) {
  QV_CHECK(sln_ok(HW, C), "spatial LayerNorm: map of %d x %d elements not covered (HW * C in {2048, 4096, 8192, 16384})", HW, C);
  const int E = HW * C;
  const size_t smem = (size_t)2 * E * 4;
  const int grid = min(B, qv_num_sms() * (E <= 4096 ? 4 : (E <= 8192 ? 2 : 1)));
  if (dt == QV_F32) SLN_DISPATCH(sln_fwd_kernel, float, (const float*)x, B, HW, C, gamma, beta, eps, (float*)y, stats);
  else SLN_DISPATCH(sln_fwd_kernel, bf16, (const bf16*)x, B, HW, C, gamma, beta, eps, (bf16*)y, stats);
  QV_LAUNCH_CHECK();
  return 0;
}

int sln_bwd(cudaStream_t s, int dt, const void* x, const void* dy, int B, int HW, int C, const float* gamma, const float* stats,
            const void* resid, void* dx, float* dgamma, float* dbeta) {
  QV_CHECK(sln_ok(HW, C), "spatial LayerNorm: map of %d x %d elements not covered", HW, C);
  const int E = HW * C;
  const size_t smem = (size_t)3 * E * 4;
  const int grid = min(B, qv_num_sms() * (E <= 4096 ? 4 : 1));
  if (dt == QV_F32) SLN_DISPATCH(sln_bwd_kernel, float, (const float*)x, (const float*)dy, B, HW, C, gamma, stats, (const float*)resid, (float*)dx, dgamma, dbeta);
  else SLN_DISPATCH(sln_bwd_kernel, bf16, (const bf16*)x, (const bf16*)dy, B, HW, C, gamma, stats, (const bf16*)resid, (bf16*)dx, dgamma, dbeta);
  QV_LAUNCH_CHECK();
  return 0;
}

int layerscale_prepare(cudaStream_t s, const float* W, const float* b, const float* gamma, int N, int K, float* Ws, float* bs) {
  qv_launch(layerscale_prepare_kernel, vgrid((long)N * K), 256, 0, s, W, b, gamma, N, K, Ws, bs);
  QV_LAUNCH_CHECK();
  return 0;
}
int layerscale_finish(cudaStream_t s, const float* G, const float* gb, const float* W, const float* b, const float* gamma, int N, int K,
                      float* dW, float* db, float* dgamma) {
  qv_launch(layerscale_finish_kernel, N, 256, 0, s, G, gb, W, b, gamma, K, dW, db, dgamma);
  QV_LAUNCH_CHECK();
  return 0;
}
int scale_rows(cudaStream_t s, int dt, const void* x, long rows, int C, const float* rowscale, int rows_per_img, const void* resid,
               void* y) {
  QV_CHECK(C % 8 == 0, "scale_rows: C = %d", C);
  const long nvec = rows * (C / 8);
  if (dt == QV_F32) qv_launch(scale_rows_kernel<float>, vgrid(nvec), 256, 0, s, (const float*)x, nvec, C, rowscale, rows_per_img, (const float*)resid, (float*)y);
  else qv_launch(scale_rows_kernel<bf16>, vgrid(nvec), 256, 0, s, (const bf16*)x, nvec, C, rowscale, rows_per_img, (const bf16*)resid, (bf16*)y);
  QV_LAUNCH_CHECK();
  return 0;
}
int stem_droppath_scales(cudaStream_t s, const unsigned long long* rng, uint32_t site0, int n_sites, int B, const float* rates, float* rs) {
  QV_CHECK(n_sites <= 7, "stem_droppath_scales: %d sites", n_sites);
  DpRates r{};
  for (int i = 0; i < n_sites; ++i) r.p[i] = rates[i];
  qv_launch(stem_droppath_kernel, cdiv(n_sites * B, 256), 256, 0, s, rng, site0, r, n_sites, B, rs);
  QV_LAUNCH_CHECK();
  return 0;
}
