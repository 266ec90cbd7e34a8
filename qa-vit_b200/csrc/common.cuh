// Shared device/host helpers for the qavit_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------- NVTX ranges (header-only nvtx3; no-ops unless a profiler is attached)
#include <nvtx3/nvToolsExt.h>
struct QvRange {
  explicit QvRange(const char* name) { nvtxRangePushA(name); }
  ~QvRange() { nvtxRangePop(); }
  QvRange(const QvRange&) = delete;
  QvRange& operator=(const QvRange&) = delete;
};
#define QV_CONCAT_(a, b) a##b
#define QV_CONCAT(a, b) QV_CONCAT_(a, b)
#define QV_RANGE(name) QvRange QV_CONCAT(qv_range_, __LINE__)(name)

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

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// Every kernel of the library starts with QV_PDL_ENTRY(): it lets the NEXT kernel of the stream be scheduled right away
// (griddepcontrol.launch_dependents) and then waits until the PREVIOUS grid has completed and its writes are visible
// (griddepcontrol.wait) -- so launch latency, CTA scheduling and (where a kernel moves the wait below its set-up code) barrier
// initialisation / TMEM allocation / descriptor prefetch overlap the tail of the producer.  Ordering stays transitive because
// every kernel waits before it touches memory.  qv_launch() attaches the launch attribute; inside a stream capture it becomes
// a programmatic edge of the CUDA graph.  OFF by default (QAVIT_PDL=1 turns the attribute on; without it both instructions
// are no-ops): on the headline step graph replay measured 46.5 ms without and 47.1-47.4 ms with programmatic edges, whatever
// the order of trigger and wait (profiles/r2_pdl_ab.txt) -- the step's kernels are long enough that the graph's ordinary
// kernel-to-kernel edge is already cheaper than the dependent-launch handshake.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#ifndef QV_PDL_MODE
#define QV_PDL_MODE 0
#endif
#if QV_PDL_MODE == 0
#define QV_PDL_ENTRY() do { pdl_trigger(); pdl_wait(); } while (0)
#elif QV_PDL_MODE == 1
#define QV_PDL_ENTRY() do { pdl_wait(); pdl_trigger(); } while (0)
#else
#define QV_PDL_ENTRY() do { pdl_wait(); } while (0)
#endif
extern int g_qv_pdl;
template <typename... KArgs, typename... Args>
inline void qv_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = g_qv_pdl ? 1 : 0;
  (void)cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);   // errors surface through QV_LAUNCH_CHECK()
}

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

// 16 B vector reduction into global memory (sm_90+): one L2 atomic operation for four floats
__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// An mma.sync C fragment (this lane: row g cols 2t, 2t + 1 and row g + 8 cols 2t, 2t + 1 of an 8-column tile) added to global memory
// with ONE vector reduction per lane instead of four scalar atomics: lanes t and t ^ 1 trade halves, so an even-t lane ends up with
// row g, cols 2t .. 2t + 3 and an odd-t lane with row g + 8, cols 2t - 2 .. 2t + 1.  row_g / row_g8: the tile's row pointers (16 B
// aligned, nullptr = row not stored).  The per-warp flushes of the attention backward kernels are millions of atomics per launch onto
// a few thousand addresses; their cost scales with the number of operations, not bytes.
__device__ __forceinline__ void red_frag_v4(float* row_g, float* row_g8, const float* c, int t) {
  const bool odd = t & 1;
  const float r0 = __shfl_xor_sync(0xffffffffu, odd ? c[0] : c[2], 1), r1 = __shfl_xor_sync(0xffffffffu, odd ? c[1] : c[3], 1);
  if (!odd) { if (row_g) red_add_v4(row_g + 2 * t, c[0], c[1], r0, r1); }
  else if (row_g8) red_add_v4(row_g8 + 2 * t - 2, r0, r1, c[2], c[3]);
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
// value and derivative from one tanh
__device__ __forceinline__ void gelu_both_fast_f(float x, float& g, float& dg) {
  const float x2 = x * x;
  const float t = tanh_approx_f(x * fmaf(0.0356774081f, x2, 0.7978845608f));
  const float hx = 0.5f * x;
  g = fmaf(hx, t, hx);
  dg = fmaf(hx * fmaf(0.1070322243f, x2, 0.7978845608f), fmaf(-t, t, 1.0f), fmaf(0.5f, t, 0.5f));
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

// ---------------------------------------------------------------- counter-based dropout masks
// Philox4x32-7: 4 x 32 random bits for (key, counter); stateless, so forward and backward regenerate the same
// dropout mask from (seed, offset, site, element id) instead of storing it.  7 rounds is the smallest Crush-resistant
// variant of Salmon et al. (SC'11, table 2; 10 is the conservative default): the dependent multiply chain sits in the
// latency-bound GEMM epilogue warps, so its length is paid almost fully.  tests/dropout_masks.py mirrors the constant.
constexpr int QV_PHILOX_ROUNDS = 7;
__device__ __forceinline__ uint4 philox4x32(uint2 key, uint4 ctr) {
#pragma unroll
  for (int r = 0; r < QV_PHILOX_ROUNDS; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
  }
  return ctr;
}
// A dropout site of the quad block (nn.Dropout / SDPA dropout_p / DropPath): p == 0 means inactive.  rng = device
// {seed, offset} snapshot taken by the forward call; 16 random bits per element (p is quantised to 1/65536 and the
// keep scale uses the quantised value, so the mask stays unbiased).
struct DropP {
  float p = 0.f;
  const unsigned long long* rng = nullptr;
  uint32_t site = 0;
};
struct DropState {
  uint2 key;
  uint32_t off_lo, off_hi, thr;
  float inv;
};
__device__ __forceinline__ DropState drop_state(const DropP& d) {
  DropState s;
  const unsigned long long seed = d.rng[0], off = d.rng[1];
  s.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32) ^ d.site);
  s.off_lo = (uint32_t)off; s.off_hi = (uint32_t)(off >> 32);
  s.thr = (uint32_t)(d.p * 65536.f + 0.5f);
  s.inv = 65536.f / (65536.f - (float)s.thr);
  return s;
}
// keep scale (0 or 1 / (1 - p)) of one element
__device__ __forceinline__ float drop_keep1(const DropState& s, unsigned long long id) {
  const uint4 r = philox4x32(s.key, make_uint4((uint32_t)id, (uint32_t)(id >> 32), s.off_lo, s.off_hi ^ 0x51u));
  return (r.x & 0xFFFFu) >= s.thr ? s.inv : 0.f;
}
// keep scales of 8 consecutive elements; id8 = (index of the first element) / 8
__device__ __forceinline__ void drop_keep8(const DropState& s, unsigned long long id8, float* sc) {
  const uint4 r = philox4x32(s.key, make_uint4((uint32_t)id8, (uint32_t)(id8 >> 32), s.off_lo, s.off_hi ^ 0x58u));
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    sc[2 * i] = (w[i] & 0xFFFFu) >= s.thr ? s.inv : 0.f;
    sc[2 * i + 1] = (w[i] >> 16) >= s.thr ? s.inv : 0.f;
  }
}
// Keep bits of a [16 x 8 NT] tile held in mma C-fragment layout x[NT][4]: bit (4 n + e) <-> x[n][e].  `tile` is any id
// unique per 16-row tile within the site; forward and backward of a kernel family use the same ids.
template <int NT>
__device__ __forceinline__ unsigned long long drop_bits_c(const DropState& s, uint32_t tile, int lane) {
  unsigned long long bits = 0;
#pragma unroll
  for (int c = 0; c < (NT + 1) / 2; ++c) {
    const uint4 r = philox4x32(s.key, make_uint4(tile, (uint32_t)(lane | (c << 5)), s.off_lo, s.off_hi ^ 0x5Cu));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if ((w[i] & 0xFFFFu) >= s.thr) bits |= 1ull << (8 * c + 2 * i);
      if ((w[i] >> 16) >= s.thr) bits |= 1ull << (8 * c + 2 * i + 1);
    }
  }
  return bits;
}
template <int NT>
__device__ __forceinline__ void drop_apply_c(float (*x)[4], unsigned long long bits, float inv) {
#pragma unroll
  for (int n = 0; n < NT; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) x[n][e] *= ((bits >> (4 * n + e)) & 1ull) ? inv : 0.f;
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
