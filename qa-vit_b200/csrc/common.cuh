// Shared device/host helpers for the qavit_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------- error plumbing (C-ABI: int status + last error)
void qv_set_error(const char* fmt, ...);
#define QV_CHECK(cond, ...)                \
  do {                                     \
    if (!(cond)) {                         \
      qv_set_error(__VA_ARGS__);           \
      return 1;                            \
    }                                      \
  } while (0)
#define QV_CUDA(expr)                                                                       \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      qv_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));    \
      return 1;                                                                             \
    }                                                                                       \
  } while (0)
extern unsigned long long g_qv_launches;   // kernels launched by this library (bench.py's gpu_launches)
#define QV_LAUNCH_CHECK()            \
  do {                               \
    ++g_qv_launches;                 \
    QV_CUDA(cudaPeekAtLastError());  \
  } while (0)
#define QV_TRY(expr)          \
  do {                        \
    int _s = (expr);          \
    if (_s) return _s;        \
  } while (0)

// ---------------------------------------------------------------- typed load/store (T = float | bf16, math in fp32)
__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// 2-wide (the channel counts here are all even)
__device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 ld2(const bf16* p) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}
__device__ __forceinline__ void st2(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
__device__ __forceinline__ void st2(bf16* p, float2 v) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __float22bfloat162_rn(v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// exact-erf GELU (nn.GELU() default) and its derivative
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// GELU for the bf16-run kernels (GEMM epilogues, lateral row kernels): the tanh form on the MUFU tanh unit, 6 / 11
// instructions instead of erff's ~35.  |gelu_tanh - gelu_erf| <= 4.8e-4 and |d gelu_tanh - d gelu_erf| <= 8.7e-4
// (measured over [-8, 8]), i.e. below the rounding of the bf16 value the result is stored as; the fp32 parity
// kernels keep the exact erf form (gelu_f / gelu_grad_f).
__device__ __forceinline__ float tanh_approx_f(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_fast_f(float x) {
  const float t = tanh_approx_f(x * fmaf(0.0356774081f, x * x, 0.7978845608f));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
__device__ __forceinline__ float gelu_grad_fast_f(float x) {
  const float x2 = x * x;
  const float t = tanh_approx_f(x * fmaf(0.0356774081f, x2, 0.7978845608f));
  const float du = fmaf(0.1070322243f, x2, 0.7978845608f);
  return fmaf(0.5f * x * du, fmaf(-t, t, 1.0f), fmaf(0.5f, t, 0.5f));
}
// type-selected: exact for fp32 activations, fast for bf16 activations
template <typename T> __device__ __forceinline__ float gelu_t(float x) { return sizeof(T) == 2 ? gelu_fast_f(x) : gelu_f(x); }
template <typename T> __device__ __forceinline__ float gelu_grad_t(float x) { return sizeof(T) == 2 ? gelu_grad_fast_f(x) : gelu_grad_f(x); }

// EPL-wide (4 or 8) vector load / store of a row slice, math in fp32 (shared by the row kernels)
template <int EPL>
__device__ __forceinline__ void load_vec(const float* p, float* v) {
#pragma unroll
  for (int i = 0; i < EPL; i += 4) {
    const float4 q = *reinterpret_cast<const float4*>(p + i);
    v[i] = q.x; v[i + 1] = q.y; v[i + 2] = q.z; v[i + 3] = q.w;
  }
}
template <int EPL>
__device__ __forceinline__ void load_vec(const bf16* p, float* v) {
  if (EPL == 8) {
    const uint4 q = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  } else {
    const uint2 q = *reinterpret_cast<const uint2*>(p);
    const uint32_t w[2] = {q.x, q.y};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
}
template <int EPL>
__device__ __forceinline__ void store_vec(float* p, const float* v) {
#pragma unroll
  for (int i = 0; i < EPL; i += 4) *reinterpret_cast<float4*>(p + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
}
template <int EPL>
__device__ __forceinline__ void store_vec(bf16* p, const float* v) {
  uint32_t w[EPL / 2];
#pragma unroll
  for (int i = 0; i < EPL / 2; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  if (EPL == 8) *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  else *reinterpret_cast<uint2*>(p) = make_uint2(w[0], w[EPL / 2 - 1]);
}

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }
static inline int qv_num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}
