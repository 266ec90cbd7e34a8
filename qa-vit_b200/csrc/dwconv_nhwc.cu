// Depthwise k x k convolutions (k = 3, 5, 7; stride 1, 'same' padding, bias) on channels-last feature maps
// [B, H, W, C] -- the ConvNeXt / LMFAdapter stencils of HQAViT's lateral path (H:722, 811-812; scope row f-1).
// One CTA per (image, 64-channel slab): the slab's H x W x 64 tile sits in shared memory, a thread owns one channel
// (its k*k taps live in registers) and walks output positions, so global traffic is each element once, coalesced
// over channels.  HBM-bound: algorithmic bytes = B*H*W*C * (in + out).
#include "../../include/qavit_b200.h"
#include "kernels.h"

namespace {

constexpr int CS = 64;     // channels per CTA
constexpr int PG = 4;      // position groups (threads per channel)

template <typename T, int K, bool FLIP>
__global__ void __launch_bounds__(CS * PG) dw_fwd_kernel(const T* __restrict__ x, int B, int H, int W, int C,
                                                         const float* __restrict__ w, const float* __restrict__ bias,
                                                         T* __restrict__ y) {
  extern __shared__ float xs[];                 // [H*W][CS]
  const int b = blockIdx.x, c0 = blockIdx.y * CS, cl = threadIdx.x % CS, pg = threadIdx.x / CS, c = c0 + cl;
  const int HW = H * W;
  const bool act = c < C;
  for (int idx = threadIdx.x; idx < HW * CS; idx += CS * PG) {
    const int p = idx / CS, cc = idx % CS;
    xs[idx] = (c0 + cc < C) ? ldf(x + ((long)b * HW + p) * C + c0 + cc) : 0.f;
  }
  float wr[K * K];
#pragma unroll
  for (int t = 0; t < K * K; ++t) wr[t] = act ? w[c * K * K + (FLIP ? K * K - 1 - t : t)] : 0.f;
  const float bs = (bias && act) ? bias[c] : 0.f;
  __syncthreads();
  if (!act) return;
  for (int p = pg; p < HW; p += PG) {
    const int py = p / W, px = p % W;
    float a = bs;
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
      const int yy = py + ky - K / 2;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const int xx = px + kx - K / 2;
        if (xx < 0 || xx >= W) continue;
        a = fmaf(wr[ky * K + kx], xs[(yy * W + xx) * CS + cl], a);
      }
    }
    stf(y + ((long)b * HW + p) * C + c, a);
  }
}

// dw[c, ky, kx] += sum_{b, p} dy[b, p, c] x[b, p + (ky, kx) - K/2, c];  dbias[c] += sum dy
template <typename T, int K>
__global__ void __launch_bounds__(CS * PG) dw_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy, int B, int H,
                                                           int W, int C, float* __restrict__ dw, float* __restrict__ dbias) {
  extern __shared__ float sm[];
  const int HW = H * W;
  float* xs = sm;                    // [HW][CS]
  float* gs = sm + HW * CS;          // [HW][CS]
  const int c0 = blockIdx.y * CS, cl = threadIdx.x % CS, pg = threadIdx.x / CS, c = c0 + cl;
  float acc[K * K], ab = 0.f;
#pragma unroll
  for (int t = 0; t < K * K; ++t) acc[t] = 0.f;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < HW * CS; idx += CS * PG) {
      const int p = idx / CS, cc = idx % CS;
      const bool ok = c0 + cc < C;
      xs[idx] = ok ? ldf(x + ((long)b * HW + p) * C + c0 + cc) : 0.f;
      gs[idx] = ok ? ldf(dy + ((long)b * HW + p) * C + c0 + cc) : 0.f;
    }
    __syncthreads();
    for (int p = pg; p < HW; p += PG) {
      const int py = p / W, px = p % W;
      const float g = gs[p * CS + cl];
      ab += g;
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
        const int yy = py + ky - K / 2;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          const int xx = px + kx - K / 2;
          if (xx < 0 || xx >= W) continue;
          acc[ky * K + kx] = fmaf(g, xs[(yy * W + xx) * CS + cl], acc[ky * K + kx]);
        }
      }
    }
  }
  // reduce the PG position groups through shared memory, then one atomic per (channel, tap) per CTA
  __syncthreads();
  float* red = sm;                   // [PG][K*K + 1][CS]
#pragma unroll
  for (int t = 0; t < K * K; ++t) red[(pg * (K * K + 1) + t) * CS + cl] = acc[t];
  red[(pg * (K * K + 1) + K * K) * CS + cl] = ab;
  __syncthreads();
  for (int idx = threadIdx.x; idx < (K * K + 1) * CS; idx += CS * PG) {
    const int t = idx / CS, cc = idx % CS;
    if (c0 + cc >= C) continue;
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < PG; ++q) s += red[(q * (K * K + 1) + t) * CS + cc];
    if (t < K * K) atomicAdd(dw + (c0 + cc) * K * K + t, s);
    else if (dbias) atomicAdd(dbias + c0 + cc, s);
  }
}

template <typename T, int K>
int run_fwd(cudaStream_t s, const T* x, int B, int H, int W, int C, const float* w, const float* bias, T* y, bool flip) {
  const size_t smem = (size_t)H * W * CS * sizeof(float);
  QV_CHECK(smem <= 200 * 1024, "dwconv: %dx%d feature map too large for one shared-memory tile", H, W);
  dim3 grid(B, cdiv(C, CS));
  if (flip) {
    if (smem > 48 * 1024) QV_CUDA(cudaFuncSetAttribute(dw_fwd_kernel<T, K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dw_fwd_kernel<T, K, true><<<grid, CS * PG, smem, s>>>(x, B, H, W, C, w, bias, y);
  } else {
    if (smem > 48 * 1024) QV_CUDA(cudaFuncSetAttribute(dw_fwd_kernel<T, K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dw_fwd_kernel<T, K, false><<<grid, CS * PG, smem, s>>>(x, B, H, W, C, w, bias, y);
  }
  QV_LAUNCH_CHECK();
  return 0;
}

template <typename T, int K>
int run_wgrad(cudaStream_t s, const T* x, const T* dy, int B, int H, int W, int C, float* dw, float* dbias) {
  const size_t tile = (size_t)2 * H * W * CS * sizeof(float), red = (size_t)PG * (K * K + 1) * CS * sizeof(float);
  const size_t smem = tile > red ? tile : red;
  QV_CHECK(smem <= 200 * 1024, "dwconv wgrad: %dx%d feature map too large", H, W);
  if (smem > 48 * 1024) QV_CUDA(cudaFuncSetAttribute(dw_wgrad_kernel<T, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int cch = cdiv(C, CS);
  const int occ = max(1, min(6, (int)(200 * 1024 / (smem + 1024))));
  dim3 grid(max(1, min(B, qv_num_sms() * occ / cch)), cch);
  dw_wgrad_kernel<T, K><<<grid, CS * PG, smem, s>>>(x, dy, B, H, W, C, dw, dbias);
  QV_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int dispatch(cudaStream_t s, int op, int K, const T* a, const T* b2, int B, int H, int W, int C, const float* w,
             const float* bias, T* out, float* dw, float* dbias) {
  switch (K) {
    case 3: return op == 2 ? run_wgrad<T, 3>(s, a, b2, B, H, W, C, dw, dbias) : run_fwd<T, 3>(s, a, B, H, W, C, w, bias, out, op == 1);
    case 5: return op == 2 ? run_wgrad<T, 5>(s, a, b2, B, H, W, C, dw, dbias) : run_fwd<T, 5>(s, a, B, H, W, C, w, bias, out, op == 1);
    case 7: return op == 2 ? run_wgrad<T, 7>(s, a, b2, B, H, W, C, dw, dbias) : run_fwd<T, 7>(s, a, B, H, W, C, w, bias, out, op == 1);
  }
  qv_set_error("dwconv: kernel size %d not supported (3, 5, 7)", K);
  return 1;
}

}  // namespace

extern "C" int qavit_dwconv_forward(const void* x, int is_bf16, int B, int H, int W, int C, int K, const float* w,
                                    const float* bias, void* y, void* stream) {
  if (B <= 0) return 0;
  if (is_bf16) return dispatch<bf16>((cudaStream_t)stream, 0, K, (const bf16*)x, nullptr, B, H, W, C, w, bias, (bf16*)y, nullptr, nullptr);
  return dispatch<float>((cudaStream_t)stream, 0, K, (const float*)x, nullptr, B, H, W, C, w, bias, (float*)y, nullptr, nullptr);
}

extern "C" int qavit_dwconv_backward(const void* x, const void* dy, int is_bf16, int B, int H, int W, int C, int K,
                                     const float* w, void* dx, float* dw, float* dbias, void* stream) {
  if (B <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (is_bf16) {
    if (dx) QV_TRY(dispatch<bf16>(s, 1, K, (const bf16*)dy, nullptr, B, H, W, C, w, nullptr, (bf16*)dx, nullptr, nullptr));
    return dispatch<bf16>(s, 2, K, (const bf16*)x, (const bf16*)dy, B, H, W, C, nullptr, nullptr, nullptr, dw, dbias);
  }
  if (dx) QV_TRY(dispatch<float>(s, 1, K, (const float*)dy, nullptr, B, H, W, C, w, nullptr, (float*)dx, nullptr, nullptr));
  return dispatch<float>(s, 2, K, (const float*)x, (const float*)dy, B, H, W, C, nullptr, nullptr, nullptr, dw, dbias);
}
